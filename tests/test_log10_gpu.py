"""The final log10 on the device (csrc/phmm_finalize.cu, phmm_log10.h): glibc 2.39's log10f / log10 restated
operation for operation.  CPU: the restatement against this host's libm (sampled here; tools/check_log10f.cpp is
the exhaustive form: all 2^31 non-negative floats, > 10^9 doubles).  GPU: the engine with the device pass against
the engine with the host log10f pass (PHMM_HOST_LOG10=1) -- bit-identical doubles on every workload shape."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restatement_matches_libm_sampled(tmp_path):
    """Compiles the exhaustive checker and runs it in its (quick) sampled mode."""
    exe = str(tmp_path / "check_log10f")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fopenmp", "-mfma", "-ffp-contract=off",
                    os.path.join(ROOT, "tools", "check_log10f.cpp"), "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe, "--quick"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert '"mismatches": 0' in r.stdout and '"double_mismatches": 0' in r.stdout


@pytest.mark.gpu
def test_device_log10_is_bit_identical_to_the_host_pass(pkg):
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from __graft_entry__ import load_package\n"
        "pkg = load_package(); S = pkg.synth\n"
        "bs = [S.s3(2), S.s2(8), S.s4(2, n_reads=24, n_haps=4), S.s5_batch(64, seed=3), S.random_small(5, n_regions=20, max_reads=30, max_haps=8, max_read_len=255, max_hap_len=400)]\n"
        "with pkg.PairHMMEngine(devices=[0]) as e:\n"
        "    out = [e.compute(b) for b in bs]\n"
        "np.savez(sys.argv[1], **{f'l{i}': o.log10 for i, o in enumerate(out)}, **{f'r{i}': o.raw32 for i, o in enumerate(out)},\n"
        "         **{f'd{i}': o.raw64 for i, o in enumerate(out)}, **{f'm{i}': o.rescued for i, o in enumerate(out)})\n")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for host in ("0", "1"):
            env = dict(os.environ)
            env.pop("PHMM_HOST_LOG10", None)
            if host == "1":
                env["PHMM_HOST_LOG10"] = "1"
            path = os.path.join(d, f"out{host}.npz")
            r = subprocess.run([sys.executable, "-c", code, path], capture_output=True, text=True, env=env, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(np.load(path))
    dev, host = outs
    assert sorted(dev.files) == sorted(host.files)
    for k in dev.files:
        a, b = dev[k], host[k]
        assert a.dtype == b.dtype and a.shape == b.shape
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), k          # bit for bit, NaN-safe
