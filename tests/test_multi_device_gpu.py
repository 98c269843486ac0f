"""GPU suite, multi-device form of the engine (phmm_options.n_devices > 1): regions are sharded over the
devices inside one process, results gathered on the host.  Skipped on a 1-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_devices_equal_one(pkg, engine):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    b = pkg.synth.random_small(61, n_regions=17, max_reads=30, max_haps=8, general_gaps=False)
    one = engine.compute(b)
    with pkg.PairHMMEngine(devices=[0, 1]) as two:
        got = two.compute(b)
        assert got.stats["n_devices_used"] == 2
        assert np.array_equal(got.log10.view(np.uint64), one.log10.view(np.uint64))
        assert np.array_equal(got.rescued, one.rescued)
        big = pkg.synth.s3(8)
        a, c = engine.compute(big, want_raw=False), two.compute(big, want_raw=False)
        assert np.array_equal(a.log10.view(np.uint64), c.log10.view(np.uint64))
        t1 = two.submit(b); t2 = two.submit(big)
        assert np.array_equal(two.wait(t1).log10.view(np.uint64), one.log10.view(np.uint64))
        assert np.array_equal(two.wait(t2).log10.view(np.uint64), a.log10.view(np.uint64))
