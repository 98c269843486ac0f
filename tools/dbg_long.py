import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from __graft_entry__ import load_package
from _oracle import load_oracle
pkg = load_package(); oracle = load_oracle()
rng = np.random.default_rng(300)
alpha = np.frombuffer(b"ACGT", np.uint8)
regions = []
for reg in range(2):
    haps = [alpha[rng.integers(0, 4, int(n))] for n in (2300, 700, 40)]
    haps[1] = haps[1].copy(); haps[1][::53] = ord("N")
    haps.append(np.concatenate([haps[0][:1000], haps[0][1003:]]))
    lens = [256, 257, 300, 511, 512, 513, 777, 1024, 1500, 2047, 2048, 100, 255, 31] if reg == 0 else [256, 1025, 64, 2048]
    reads, quals = [], []
    for rl in lens:
        h = haps[0]
        o = int(rng.integers(0, len(h) - rl + 1)); r = h[o:o + rl].copy()
        m = rng.random(rl) < 0.01; r[m] = alpha[rng.integers(0, 4, int(m.sum()))]
        reads.append(r); quals.append((33 + rng.integers(20, 42, rl)).astype(np.uint8))
        for _ in range(3): rng.integers(30, 50, rl)
    regions.append((reads, quals, haps))
b = pkg.Batch.from_regions(regions)
want = oracle.batch(b, threads=16)
with pkg.PairHMMEngine(devices=[0], exact_fp32=1) as eng:
    got = eng.compute(b)
i = 0
for g, (reads, quals, haps) in enumerate(regions):
    for r in reads:
        for h in haps:
            gb, wb = got.raw32[i:i+1].view(np.uint32)[0], want["raw32"][i:i+1].view(np.uint32)[0]
            g6, w6 = got.raw64[i:i+1].view(np.uint64)[0], want["raw64"][i:i+1].view(np.uint64)[0] if "raw64" in want else 0
            if gb != wb or (want["rescued"][i] and "raw64" in want and g6 != w6):
                print(f"region {g} R={len(r)} H={len(h)} resc={want['rescued'][i]} raw32 got={got.raw32[i]:.9e} ({gb:#x}) want={want['raw32'][i]:.9e} ({wb:#x}) raw64 {got.raw64[i]:.17e} {want.get('raw64', [0]*(i+1))[i]:.17e}")
            i += 1
print("keys", list(want.keys()))
