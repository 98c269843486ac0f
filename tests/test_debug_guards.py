"""compute-sanitizer is closed on the GPU pool this repo is measured on (profiles/r02_sanitize_memcheck_refused.txt),
so the memcheck / initcheck questions are asked with the library's own debug aids (csrc/phmm_engine.cu: DeviceBuf):

  out-of-bounds writes   PHMM_DEBUG_GUARD=1 puts every device buffer between two 4 KB guard zones of 0xA5;
                         phmm_debug_check() counts overwritten guard bytes after a pass over every kernel family
                         (tools/sanitize_target.py: all shapes, modes, tiers, work lists, long reads, genotype
                         reduction) -- must be 0;
  uninitialised reads    the same pass with the buffers pre-filled with 0x00, with 0xFF and with 0x7F (NaN patterns
                         included) must give BIT-IDENTICAL results: a kernel that read device memory nobody wrote
                         would see different bytes in each run.
Results are also checked against the oracle inside the target, so a wrong-but-stable answer does not pass."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_guard_zones_intact_and_results_independent_of_poison(tmp_path):
    outs = []
    for poison in ("0x00", "0xff", "0x7f"):
        path = str(tmp_path / f"dump_{poison}.npz")
        env = dict(os.environ, PHMM_DEBUG_GUARD="1", PHMM_POISON=poison)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_target.py"), "--dump", path],
                           capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert "guard bytes overwritten: [0, 0, 0, 0]" in r.stdout and "SANITIZE TARGET DONE" in r.stdout
        outs.append(np.load(path))
    keys = sorted(outs[0].files)
    assert len(keys) > 30
    for other in outs[1:]:
        assert sorted(other.files) == keys
        for k in keys:
            assert np.array_equal(outs[0][k], other[k]), k
