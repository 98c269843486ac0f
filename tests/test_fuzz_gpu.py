"""A short run of tools/fuzz_gpu.py: random ragged batches, random engine configuration (workers, FP64-first, exact),
random variant sites -- everything against the oracle.  (profiles/r02_fuzz.json is a 250 s run: 19 787 batches,
11.3 M pairs, 452 628 sites, 0 disagreements.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_fuzz_against_the_oracle_for_20_seconds():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_gpu.py"), "20", "777"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["disagreements"] == 0 and out["batches"] > 100 and out["sites"] > 1000
