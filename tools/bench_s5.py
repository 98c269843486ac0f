"""S5 (BASELINE configs[4]): ragged active-region stream in the reference's windowing; e2e through submit/wait."""
import sys, os, time
from collections import deque
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wpb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
t0 = time.perf_counter()
batches = list(pkg.synth.s5_stream(n_win, windows_per_batch=wpb))
print(f"generated {len(batches)} batches of {wpb} windows in {time.perf_counter()-t0:.1f} s; "
      f"pairs/batch {batches[0].n_pairs}, cells/batch {batches[0].n_cells:.3e}, reads/region mean {batches[0].reads_per_region.mean():.1f}, "
      f"haps/region mean {batches[0].haps_per_region.mean():.1f}, read len {np.diff(batches[0].read_off).min()}..{np.diff(batches[0].read_off).max()}")
with pkg.PairHMMEngine(devices=[0], pipeline_depth=4, host_threads=4) as eng:
    for b in batches[:4]: eng.compute(b, want_raw=False)
    st = eng.stage(batches[0]); eng.run_staged(st, 2); ms, nl = eng.run_staged(st, 5)
    print(f"kernels only: {ms:.3f} ms/batch, {batches[0].n_cells/ms/1e6:.1f} GCUPS, launches {nl}")
    eng.free_staged(st)
    cells = sum(b.n_cells for b in batches)
    t0 = time.perf_counter(); q = deque(); i = 0; done = 0; resc = 0
    while done < len(batches):
        while len(q) < 4 and i < len(batches): q.append(eng.submit(batches[i])); i += 1
        r = eng.wait(q.popleft()); resc += r.stats["n_rescued"]; done += 1
    dt = time.perf_counter() - t0
    print(f"e2e stream: {cells:.3e} cells in {dt*1e3:.1f} ms = {cells/dt/1e9:.1f} GCUPS, rescued {resc} of {sum(b.n_pairs for b in batches)} pairs")
