"""Where does the end-to-end time go?  Times submit / wait separately (development aid)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
nb, regions, steps = 6, 128, 24
ht = int(sys.argv[1]) if len(sys.argv) > 1 else 4
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batches = [pkg.synth.s3(regions, seed=1003 + i) for i in range(nb)]
with pkg.PairHMMEngine(devices=[0], pipeline_depth=depth, host_threads=ht) as eng:
    res = [pkg.Result(b.n_pairs, want_raw=False) for b in batches[:2]]
    if depth > 2:
        from collections import deque
        for w in range(4): eng.compute(batches[w % nb], want_raw=False)
        q = deque(); t_all = time.perf_counter(); done = 0; s_i = 0
        while done < steps:
            while len(q) < depth and s_i < steps:
                q.append(eng.submit(batches[s_i % nb])); s_i += 1
            eng.wait(q.popleft(), result=res[done % 2]); done += 1
        tot = time.perf_counter() - t_all
        print(f"depth={depth} host_threads={ht} per step: total {1e3*tot/steps:.2f} ms GCUPS {batches[0].n_cells*steps/tot/1e9:.0f}")
        sys.exit(0)
    for w in range(4): eng.compute(batches[w % nb], want_raw=False)
    ts, tw, km = [], [], []
    t_all = time.perf_counter()
    t0 = time.perf_counter(); tk = eng.submit(batches[0]); ts.append(time.perf_counter() - t0)
    for s in range(1, steps + 1):
        nxt = None
        if s < steps:
            t0 = time.perf_counter(); nxt = eng.submit(batches[s % nb]); ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); r = eng.wait(tk, result=res[s % 2]); tw.append(time.perf_counter() - t0)
        km.append(r.stats["kernel_ms"])
        tk = nxt
    tot = time.perf_counter() - t_all
    print(f"host_threads={ht} per step: total {1e3*tot/steps:.2f} ms | submit {1e3*np.mean(ts):.2f} ms | wait {1e3*np.mean(tw):.2f} ms | kernel {np.mean(km):.2f} ms | GCUPS {batches[0].n_cells*steps/tot/1e9:.0f}")
    # serial reference: compute() one by one
    t0 = time.perf_counter()
    for s in range(6): r = eng.compute(batches[s % nb], want_raw=False)
    print(f"   serial compute: {1e3*(time.perf_counter()-t0)/6:.2f} ms per batch; stats total_ms {r.stats['total_ms']:.2f} kernel_ms {r.stats['kernel_ms']:.2f}")
