"""GPU suite, multi-device form of the engine (phmm_options.n_devices > 1): ONE host process, a worker per
device, the regions of every batch sharded over the devices by cell count, results gathered on the host
(the serial loop it parallelises is haplotypecaller.hpp:138-152).

A device ordinal may be listed more than once -- every entry is an independent worker with its own streams,
tables and memory pool -- so the whole scheduler (split, per-device pack / plan / launch, output offsets,
gather, error paths) is exercised on a ONE-GPU box with devices=[0, 0] / [0, 0, 0]; on a multi-GPU box the
same tests also run over distinct GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], list(range(min(n, 8)))]
    return lists


def test_sharded_engine_equals_one_device(pkg, engine):
    b = pkg.synth.random_small(61, n_regions=17, max_reads=30, max_haps=8, general_gaps=False)
    ragged = pkg.synth.s5_batch(96, seed=5)
    big = pkg.synth.s3(8)
    one = engine.compute(b)
    one_ragged = engine.compute(ragged)
    one_big = engine.compute(big, want_raw=False)
    for devs in _device_lists():
        with pkg.PairHMMEngine(devices=devs, pipeline_depth=3, host_threads=2) as eng:
            got = eng.compute(b)
            assert got.stats["n_devices_used"] == len(devs), devs
            assert np.array_equal(got.log10.view(np.uint64), one.log10.view(np.uint64)), devs
            assert np.array_equal(got.rescued, one.rescued)
            gr = eng.compute(ragged)
            assert np.array_equal(gr.log10.view(np.uint64), one_ragged.log10.view(np.uint64)), devs
            assert gr.stats["n_cells"] == ragged.n_cells and gr.stats["n_pairs"] == ragged.n_pairs
            # several tickets in flight, waited out of order, pinned and pageable inputs
            t1 = eng.submit(b); t2 = eng.submit(big); t3 = eng.submit(ragged)
            assert np.array_equal(eng.wait(t2).log10.view(np.uint64), one_big.log10.view(np.uint64))
            assert np.array_equal(eng.wait(t1).log10.view(np.uint64), one.log10.view(np.uint64))
            assert np.array_equal(eng.wait(t3).log10.view(np.uint64), one_ragged.log10.view(np.uint64))
            ragged.pin()
            try:
                assert np.array_equal(eng.compute(ragged).log10.view(np.uint64), one_ragged.log10.view(np.uint64))
            finally:
                ragged.unpin()
            # fewer regions than devices: the empty shares are skipped
            tiny = b.slice_regions(0, 1)
            gt = eng.compute(tiny)
            assert gt.stats["n_devices_used"] == 1
            assert np.array_equal(gt.log10.view(np.uint64), one.log10[:tiny.n_pairs].view(np.uint64))


def test_sharded_engine_vs_oracle(pkg, oracle):
    """The shares of a sharded batch against the CPU oracle directly (not only against the 1-device engine)."""
    b = pkg.synth.s5_batch(48, seed=11)
    want = oracle.batch(b, threads=16)
    with pkg.PairHMMEngine(devices=[0, 0, 0, 0], host_threads=2) as eng:
        got = eng.compute(b)
    resc = want["rescued"].astype(bool)
    assert np.array_equal(got.rescued.astype(bool), resc)
    assert np.abs(got.log10[~resc] - want["log10"][~resc]).max() <= 1e-4
    if resc.any():
        assert np.abs(got.log10[resc] - want["log10"][resc]).max() <= 1e-9


def test_submit_failure_on_one_device_drains_the_others(pkg):
    """A share that fails to pack (injected) fails the submit; the sibling devices' uploads and launches are
    waited out before their slots are reused, and the engine keeps working (phmm_engine.cu: drain_parts)."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from __graft_entry__ import load_package\n"
        "pkg = load_package()\n"
        "b = pkg.synth.s3(8)\n"
        "cut = pkg.shard_bounds(b.region_cells(), 2)\n"
        "import os; os.environ['PHMM_FAULT_PACK'] = str(cut[1])\n"
        "with pkg.PairHMMEngine(devices=[0, 0], pipeline_depth=2) as eng:\n"
        "    try:\n"
        "        eng.compute(b); print('NOFAIL')\n"
        "    except pkg.PhmmError as ex:\n"
        "        print('FAILED', ex.code, 'injected' in str(ex))\n"
        "    want = None\n"
        "    for _ in range(5):\n"           # every slot of both workers is usable again
        "        got = eng.compute(b, want_raw=False).log10\n"
        "        assert want is None or np.array_equal(got, want); want = got\n"
        "    print('HEALTHY', int(np.isfinite(want).sum()))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "FAILED 3 True" in r.stdout and "HEALTHY" in r.stdout, r.stdout
