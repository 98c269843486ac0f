"""Where does the in-process multi-GPU time go? (development aid)"""
import os, sys, time
from collections import deque
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package()
b = next(pkg.synth.s5_stream(8192, windows_per_batch=8192))
batches = [b] * 12
n = min(8, torch.cuda.device_count())
with pkg.PairHMMEngine(devices=list(range(n)), pipeline_depth=4, host_threads=4) as eng:
    for x in batches[:4]: eng.compute(x, want_raw=False)
    results = [pkg.Result(b.n_pairs, want_raw=False) for _ in range(2)]
    ts, tw = [], []
    q = deque(); i = done = 0
    t0 = time.perf_counter()
    while done < len(batches):
        while len(q) < 3 and i < len(batches):
            a = time.perf_counter(); q.append(eng.submit(batches[i])); ts.append(time.perf_counter() - a); i += 1
        a = time.perf_counter(); r = eng.wait(q.popleft(), result=results[done % 2]); tw.append(time.perf_counter() - a); done += 1
    dt = time.perf_counter() - t0
    print(f"devices {n}: {1e3*dt/len(batches):.2f} ms/batch, submit {1e3*np.mean(ts):.2f} ms, wait {1e3*np.mean(tw):.2f} ms, kernel_ms(max dev) {r.stats['kernel_ms']:.2f}, GCUPS {b.n_cells*len(batches)/dt/1e9:.0f}")
