"""Quick kernel-only timing of the synthetic shapes (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
which = sys.argv[1:] or ["s3", "s2", "s3g", "s4"]
mk = {"s3": lambda: pkg.synth.s3(64), "s2": lambda: pkg.synth.s2(256), "s3g": lambda: pkg.synth.s3(64, general_gaps=True),
      "s4": lambda: pkg.synth.s4(8)}
with pkg.PairHMMEngine(devices=[0]) as eng:
    for name in which:
        b = mk[name]()
        st = eng.stage(b)
        eng.run_staged(st, 2)
        best = 1e9
        for _ in range(3):
            ms, n = eng.run_staged(st, 5); best = min(best, ms)
        print(f"{name:5s} FORCE_GROUP={os.environ.get('PHMM_FORCE_GROUP','-'):3s} cells={b.n_cells:.3e} ms/iter={best:.3f} GCUPS={b.n_cells/best/1e6:.1f} launches={n}", flush=True)
        eng.free_staged(st)
