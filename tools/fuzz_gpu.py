"""Randomised stress of the whole engine against the oracle for a time budget (default 150 s): random ragged batches
(tiny to 255-base reads, occasional long reads, N bases, per-base or constant gap penalties, hopeless reads that are
rescued / filtered), random engine (one worker or several on the same GPU, FP64-first forced or automatic, exact or
fast), random variant sites for the device-side genotype reduction.  Prints one JSON summary; exits non-zero on the
first disagreement (with the seed that reproduces it)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
from _oracle import load_oracle  # noqa: E402
from test_genotype import expected_gl, random_sites  # noqa: E402

pkg = load_package()
oracle = load_oracle()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 150.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
engines = {}


def engine(devs, first, exact):
    key = (tuple(devs), first, exact)
    if key not in engines:
        engines[key] = pkg.PairHMMEngine(devices=list(devs), pipeline_depth=3, host_threads=2, fp64_first=first, exact_fp32=exact)
    return engines[key]


t0 = time.time()
n_batches = n_pairs = n_sites = n_rescued = 0
while time.time() - t0 < budget:
    seed = int(rng.integers(1, 2**31))
    r = np.random.default_rng(seed)
    kw = dict(n_regions=int(r.integers(1, 30)), max_reads=int(r.integers(1, 40)), max_haps=int(r.integers(1, 12)),
              max_read_len=int(r.choice([8, 40, 100, 160, 255])), max_hap_len=int(r.choice([5, 60, 200, 500])),
              general_gaps=bool(r.random() < 0.3), n_frac=float(r.choice([0.0, 0.03, 0.2])), lower_frac=float(r.choice([0.0, 0.1])))
    b = pkg.synth.random_small(seed, **kw)
    if r.random() < 0.15:                                          # a dense-rescue batch in between: flips the automatic order
        b = pkg.Batch.concat([b, pkg.synth.s4(1, n_reads=int(r.integers(2, 12)), n_haps=int(r.integers(1, 4)), hap_lo=200, hap_hi=320,
                                              seed=seed)]) if not kw["general_gaps"] else b
    devs = [0] * int(r.choice([1, 1, 2, 3]))
    first = int(r.choice([0, 0, 2]))
    exact = bool(r.random() < 0.25)
    eng = engine(devs, first, exact)
    want = oracle.batch(b, threads=8)
    got = eng.compute(b)
    resc = want["rescued"].astype(bool)
    what = f"seed {seed} kw {kw} devs {devs} fp64_first {first} exact {exact}"
    if not np.array_equal(got.rescued.astype(bool), resc):
        print("RESCUE DECISIONS DIFFER:", what); sys.exit(1)
    with np.errstate(invalid="ignore"):           # -inf on both sides compares equal; its difference is never used
        d = np.where(got.log10 == want["log10"], 0.0, np.abs(got.log10 - want["log10"]))
    d = np.nan_to_num(d, nan=np.inf)
    if (d[~resc] > 1e-4).any() or (d[resc] > 1e-9).any():
        print("PARITY:", what, float(d[~resc].max(initial=0)), float(d[resc].max(initial=0))); sys.exit(1)
    if exact and not np.array_equal(got.raw32.view(np.uint32), want["raw32"].view(np.uint32)):
        print("EXACT RAW BITS:", what); sys.exit(1)
    # device-side genotype reduction against the oracle fed with this engine's own matrix
    sites, per_site = random_sites(pkg, b, seed, with_overlap=bool(r.random() < 0.7), max_sites=3)
    if per_site:
        exp, exp_n, _, _ = expected_gl(pkg, oracle, b, per_site, got.log10)
        gl = eng.compute_gl(b, sites)
        for k in range(len(per_site)):
            if gl.site(k).view(np.uint64).tolist() != exp[k].view(np.uint64).tolist() or gl.site_n_reads[k] != exp_n[k]:
                print("GENOTYPE LIKELIHOODS:", what, "site", k); sys.exit(1)
        n_sites += len(per_site)
    n_batches += 1; n_pairs += b.n_pairs; n_rescued += int(resc.sum())
for e in engines.values():
    e.close()
print(json.dumps({"seconds": round(time.time() - t0, 1), "batches": n_batches, "pairs": n_pairs, "rescued_pairs": n_rescued,
                  "sites": n_sites, "engines": len(engines), "disagreements": 0}))
