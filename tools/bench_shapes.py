"""Every BASELINE.json synthetic config on one GPU (supplement to bench.py, which times S3 only):
kernels-only GCUPS (inputs resident in HBM, CUDA events in-library) and end-to-end GCUPS through
phmm_submit / phmm_wait with host buffers (4 batches in flight).  One JSON object on stdout."""
import json, os, sys, time
from collections import deque
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
S = pkg.synth
PEAK = 148 * 128 * 1.965 / 8          # FP32 CUDA-core roofline, GCUPS
cfgs = {
    "S2 100x300, 64x8, 1024 regions (configs[1])": lambda i: S.s2(1024, seed=1002 + i),
    "S3 150x500, 256x16, 128 regions (configs[2])": lambda i: S.s3(128, seed=1003 + i),
    "S3 with per-base gap penalties (general mode)": lambda i: S.s3(128, general_gaps=True, seed=1003 + i),
    "S4 150-250 x 600-1000, 128x16, low-quality tails, all pairs redone in FP64, 64 regions (configs[3])": lambda i: S.s4(64, seed=1004 + i),
}
EXACT = "--exact" in sys.argv          # exact_fp32 engine: unfused arithmetic, raw FP32 sums bit-identical to the reference
out = {"peak_fp32_gcups": round(PEAK, 1), "engine": "exact_fp32" if EXACT else "default (FMA-contracted)", "rows": []}
with pkg.PairHMMEngine(devices=[0], pipeline_depth=4, host_threads=4, exact_fp32=EXACT) as eng:
    def measure(name, batches):
        for b in batches[:4]: eng.compute(b, want_raw=False)
        st = eng.stage(batches[0]); eng.run_staged(st, 2)
        ms = min(eng.run_staged(st, 5)[0] for _ in range(3)); nl = eng.run_staged(st, 1)[1]
        eng.free_staged(st)
        cells = sum(b.n_cells for b in batches); q = deque(); i = done = resc = 0
        t0 = time.perf_counter()
        while done < len(batches):
            while len(q) < 4 and i < len(batches): q.append(eng.submit(batches[i])); i += 1
            r = eng.wait(q.popleft()); resc += r.stats["n_rescued"]; done += 1
        dt = time.perf_counter() - t0
        k = batches[0].n_cells / ms / 1e6
        out["rows"].append({"config": name, "pairs_per_batch": int(batches[0].n_pairs), "cells_per_batch": int(batches[0].n_cells),
                            "kernel_ms_per_batch": round(ms, 3), "kernels_only_gcups": round(k, 1), "frac_of_fp32_roofline": round(k / PEAK, 3),
                            "launches_per_batch": nl, "e2e_gcups": round(cells / dt / 1e9, 1), "e2e_batches": len(batches),
                            "rescued_frac": round(resc / sum(b.n_pairs for b in batches), 4)})
        print(out["rows"][-1], file=sys.stderr)
    for name, mk in cfgs.items():
        measure(name, [mk(i) for i in range(8)])
    measure("S5 ragged window stream, 30x coverage, 1024 windows per batch (configs[4], one GPU's share)",
            list(S.s5_stream(8192, windows_per_batch=1024)))
print(json.dumps(out))
