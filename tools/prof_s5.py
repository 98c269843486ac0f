"""Profiling target: one staged S5 batch (ragged window stream), forward kernels only."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = next(pkg.synth.s5_stream(n, windows_per_batch=n))
with pkg.PairHMMEngine(devices=[0]) as eng:
    eng.compute(b, want_raw=False)
    st = eng.stage(b)
    eng.run_staged(st, 1)
    ms, nl = eng.run_staged(st, 1)
    print(f"s5 x{n}: cells={b.n_cells:.3e} ms/iter={ms:.3f} GCUPS={b.n_cells/ms/1e6:.1f} launches/iter={nl}")
    eng.free_staged(st)
