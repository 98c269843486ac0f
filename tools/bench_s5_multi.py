"""BASELINE configs[4]: the ragged active-region stream, region-sharded over the GPUs of one box INSIDE one
process (phmm_options.n_devices = N: one worker thread + host pool + stream ring per device, regions of
every batch cut by cell count, results gathered on the host, no collective).  End-to-end GCUPS through
phmm_submit / phmm_wait for N = 1, 2, 4, 8 (as many as are visible).  One JSON object on stdout."""
import json, os, sys, time
from collections import deque
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
pkg = load_package()
n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
wpb = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # the distinct batches are streamed `reps` times over
batches = list(pkg.synth.s5_stream(n_win, windows_per_batch=wpb)) * reps
pinned = len(sys.argv) > 4 and sys.argv[4] == "pinned"     # PHMM_BATCH_PINNED_INPUTS: no staging copy of the byte arrays
if pinned:
    for b in {id(x): x for x in batches}.values(): b.pin()
cells = sum(b.n_cells for b in batches)
out = {"pinned_inputs": pinned, "windows": n_win * reps, "distinct_windows": n_win, "windows_per_batch": wpb, "batches": len(batches), "cells": int(cells),
       "pairs": int(sum(b.n_pairs for b in batches)), "host_cores": os.cpu_count(), "rows": []}
for n in (1, 2, 4, 8):
    if n > torch.cuda.device_count(): break
    with pkg.PairHMMEngine(devices=list(range(n)), pipeline_depth=4, host_threads=4) as eng:
        for b in batches[:4]: eng.compute(b, want_raw=False)          # grows every slot's buffers
        best = None
        results = [pkg.Result(max(b.n_pairs for b in batches), want_raw=False) for _ in range(2)]   # caller-owned, reused
        for rep in range(3):
            q = deque(); i = done = 0
            t0 = time.perf_counter()
            while done < len(batches):
                while len(q) < 3 and i < len(batches): q.append(eng.submit(batches[i])); i += 1
                r = eng.wait(q.popleft(), result=results[done % 2]); done += 1
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["rows"].append({"n_devices": n, "e2e_ms": round(best * 1e3, 2), "e2e_gcups": round(cells / best / 1e9, 1),
                            "devices_used_last_batch": int(r.stats["n_devices_used"])})
        print(out["rows"][-1], file=sys.stderr)
print(json.dumps(out))
