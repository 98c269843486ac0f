"""Pins the oracle: a time-boxed campaign of the plain-C restatements under oracle/ against the reference's own code
compiled in place (oracle/_ref, built by oracle/Makefile from /root/reference; this container only -- the tool needs
that build and is not part of the GPU-box runs).  Random ragged batches through both PairHMM implementations (raw FP32
/ FP64 sums, rescue decisions, log10 doubles: bit for bit), random variant sites through both genotype reductions,
random sequence pairs through both Smith-Waterman aligners.  Prints one JSON summary; exits non-zero on the first
disagreement with the seed that reproduces it.   usage: oracle_vs_reference.py [seconds] [master seed]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
from _oracle import load_oracle, load_ref  # noqa: E402
from test_genotype import gl_case  # noqa: E402
from test_sw import PARAMS, _ARGS, _oracle_align, sw_cases  # noqa: E402

pkg = load_package()
oracle, ref = load_oracle(), load_ref()
if ref is None:
    sys.exit("oracle/_ref/libref_pairhmm.so is not built (make -C oracle, with /root/reference present)")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
threads = os.cpu_count() or 1
ref.lib.ref_sw_align.argtypes = _ARGS
oracle.lib.sw_oracle_align.argtypes = _ARGS

t0 = time.time()
n_batches = n_pairs = n_rescued = n_gl = n_sw = 0
while time.time() - t0 < budget:
    seed = int(rng.integers(1, 2**31))
    r = np.random.default_rng(seed)
    kw = dict(n_regions=int(r.integers(1, 12)), max_reads=int(r.integers(1, 40)), max_haps=int(r.integers(1, 12)),
              max_read_len=int(r.choice([8, 40, 100, 160, 255])), max_hap_len=int(r.choice([5, 60, 200, 500])),
              general_gaps=bool(r.random() < 0.4), n_frac=float(r.choice([0.0, 0.03, 0.2])), lower_frac=float(r.choice([0.0, 0.1])))
    b = pkg.synth.random_small(seed, **kw)
    if r.random() < 0.15 and not kw["general_gaps"]:               # long pairs that all underflow FP32
        b = pkg.Batch.concat([b, pkg.synth.s4(1, n_reads=int(r.integers(2, 6)), n_haps=int(r.integers(1, 3)), hap_lo=200, hap_hi=320, seed=seed)])
    a, c = oracle.batch(b, threads=threads), ref.batch(b, threads=threads)
    for key, ty in (("raw32", np.uint32), ("raw64", np.uint64), ("log10", np.uint64)):
        if not np.array_equal(a[key].view(ty), c[key].view(ty)):
            print("PAIRHMM", key, "seed", seed, kw); sys.exit(1)
    if not np.array_equal(a["rescued"], c["rescued"]):
        print("PAIRHMM rescue decisions, seed", seed, kw); sys.exit(1)
    n_batches += 1; n_pairs += b.n_pairs; n_rescued += int(a["rescued"].sum())
    for k in range(8):                                             # genotype reduction (genotyper.hpp:271-327)
        case = gl_case(seed % 1000003 + k)
        ga, na = oracle.genotype_likelihoods(*case)
        gb, nb = ref.genotype_likelihoods(*case)
        if na != nb or not np.array_equal(ga.view(np.uint64), gb.view(np.uint64)):
            print("GENOTYPE LIKELIHOODS seed", seed % 1000003 + k); sys.exit(1)
        n_gl += 1
    for k, (rs, al) in enumerate(sw_cases(seed % 1000003, 6, max_len=int(r.choice([60, 200, 400])))):   # intel_smithwaterman.hpp:29-44
        p = PARAMS[k % 4]
        if _oracle_align(oracle.lib, "sw_oracle_align", rs, al, p) != _oracle_align(ref.lib, "ref_sw_align", rs, al, p):
            print("SMITH-WATERMAN seed", seed % 1000003, "case", k); sys.exit(1)
        n_sw += 1
print(json.dumps({"seconds": round(time.time() - t0, 1), "threads": threads, "pairhmm_batches": n_batches, "pairs": n_pairs,
                  "rescued_pairs": n_rescued, "genotype_sites": n_gl, "sw_alignments": n_sw, "disagreements": 0}))
