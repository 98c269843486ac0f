"""Profiling target: one staged S3 batch (150x500, 256x16 per region), forward kernels only."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
kind = sys.argv[2] if len(sys.argv) > 2 else "s3"
b = {"s3": lambda: pkg.synth.s3(n), "s3g": lambda: pkg.synth.s3(n, general_gaps=True),
     "s2": lambda: pkg.synth.s2(n), "s4": lambda: pkg.synth.s4(n)}[kind]()
with pkg.PairHMMEngine(devices=[0]) as eng:
    eng.compute(b, want_raw=False)          # as in a stream: the order of the precision passes follows the previous batch
    st = eng.stage(b)
    eng.run_staged(st, 1)
    ms, nl = eng.run_staged(st, 3)
    print(f"{kind} x{n}: cells={b.n_cells:.3e} ms/iter={ms:.3f} GCUPS={b.n_cells/ms/1e6:.1f} launches/iter={nl}")
    eng.free_staged(st)
