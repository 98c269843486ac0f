"""Which pairs of the all-lengths test disagree with the oracle? (development aid)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from __graft_entry__ import load_package
from _oracle import load_oracle
pkg = load_package(); oracle = load_oracle()
rng = np.random.default_rng(77)
alpha = np.frombuffer(b"ACGT", np.uint8)
regions = []
for reg in range(2):
    haps = [alpha[rng.integers(0, 4, int(n))] for n in (311, 97, 5)]
    lens = np.arange(1 + reg, 256, 2); rng.shuffle(lens)
    reads, quals = [], []
    for rl in lens:
        rl = int(rl); h = haps[0]
        o = int(rng.integers(0, len(h) - rl + 1)); r = h[o:o + rl].copy()
        m = rng.random(rl) < 0.03; r[m] = alpha[rng.integers(0, 4, int(m.sum()))]
        reads.append(r); quals.append((33 + rng.integers(2, 42, rl)).astype(np.uint8))
    regions.append((reads, quals, haps))
b = pkg.Batch.from_regions(regions)
want = oracle.batch(b, threads=16)
for exact in (0, 1):
    with pkg.PairHMMEngine(devices=[0], exact_fp32=exact) as eng:
        got = eng.compute(b)
    i = 0
    for g, (reads, quals, haps) in enumerate(regions):
        for r in reads:
            for hi, h in enumerate(haps):
                a, w = got.log10[i], want["log10"][i]
                bad = not (a == w or abs(a - w) <= (1e-9 if want["rescued"][i] else 1e-5))
                if exact: bad = bad or (got.raw32[i:i+1].view(np.uint32)[0] != want["raw32"][i:i+1].view(np.uint32)[0])
                if bad: print(f"exact={exact} region {g} R={len(r)} H={len(h)} resc={want['rescued'][i]} got={a:.8f} want={w:.8f} raw32 {got.raw32[i]:.6e} {want['raw32'][i]:.6e}")
                i += 1
