"""How much do the hand-overs between haplotypes cost?  Same cells, long vs short haplotypes (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
with pkg.PairHMMEngine(devices=[0]) as eng:
    for H, nh in ((250, 32), (500, 16), (1000, 8), (2000, 4), (4000, 2)):
        b = pkg.synth.fixed_shape(n_regions=64, read_len=150, hap_len=H, n_reads=256, n_haps=nh, seed=5)
        info, _ = pkg.plan(b)
        st = eng.stage(b); eng.run_staged(st, 2)
        ms = min(eng.run_staged_ex(st, 5)[1] for _ in range(3))
        print(f"H={H} nh={nh} hpj={info['haps_per_job']} cells={b.n_cells:.3e} fp32 kernel ms={ms:.3f} GCUPS={b.n_cells/ms/1e6:.1f}", flush=True)
        eng.free_staged(st)
