"""Sanitizer target: runs under compute-sanitizer where that tool is available (tools/sanitize.sh) and, where it is
not (it is closed on this pool: profiles/r02_sanitize_memcheck_refused.txt), under the library's own guard zones and
poison fill (PHMM_DEBUG_GUARD=1, PHMM_POISON=<byte>; tests/test_debug_guards.py):  `--dump out.npz` stores every
result so that runs under different poison bytes can be compared bit for bit.  One small pass over EVERY kernel family of libphmm_b200.so -- the
ragged / lane-aligned / packed FP32 kernels in all three gap modes, the FP64 redo through its work list, the
FP64-first order, the flush-exact tier, the long-read kernel, the device log10 + genotype reduction, the
Smith-Waterman kernel -- with results checked against the oracle, so a sanitizer run is also a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
from _oracle import load_oracle  # noqa: E402
from _sw_cases import sw_cases  # noqa: E402

pkg = load_package()
S = pkg.synth
oracle = load_oracle()


DUMP = {}


def check(got, b, what):
    DUMP[f"{len(DUMP):03d} {what}"] = np.concatenate([got.log10.view(np.uint8), got.raw32.view(np.uint8), got.raw64.view(np.uint8), got.rescued])
    want = oracle.batch(b, threads=8)
    resc = want["rescued"].astype(bool)
    assert np.array_equal(got.rescued.astype(bool), resc), what
    e32 = np.abs(got.log10[~resc] - want["log10"][~resc]).max() if (~resc).any() else 0.0
    d = got.log10[resc] - want["log10"][resc]
    e64 = np.abs(d[np.isfinite(d)]).max() if np.isfinite(d).any() else 0.0
    assert e32 <= 1e-4 and e64 <= 1e-9, (what, e32, e64)
    print(f"ok  {what}: {b.n_pairs} pairs, {int(resc.sum())} rescued, launches {got.stats['kernel_launches']}", flush=True)


rng = np.random.default_rng(3)
alpha = np.frombuffer(b"ACGT", np.uint8)


def lengths_batch(lens, general):
    hap = alpha[rng.integers(0, 4, 180)]
    reads, quals, gi, gd, gc = [], [], [], [], []
    for rl in lens:
        o = int(rng.integers(0, max(1, len(hap) - rl + 1)))
        r = np.resize(hap[o:o + rl], rl).copy()
        reads.append(r); quals.append((33 + rng.integers(2, 42, rl)).astype(np.uint8))
        gi.append((33 + rng.integers(20, 50, rl)).astype(np.uint8)); gd.append((33 + rng.integers(20, 50, rl)).astype(np.uint8))
        gc.append((33 + rng.integers(5, 25, rl)).astype(np.uint8))
    haps = [hap, hap[:41], hap[:5]]
    return pkg.Batch.from_regions([(reads, quals, haps, gi, gd, gc) if general else (reads, quals, haps)])


batches = [
    ("ragged, per-base gaps (MODE 0)", S.random_small(1, n_regions=4, max_reads=8, max_haps=4)),
    ("ragged, constant gaps (MODE 2)", S.random_small(2, n_regions=4, max_reads=8, max_haps=4, general_gaps=False)),
    ("every K of G=16 and G=32, odd K too (MODE 2)", lengths_batch([7, 15, 16, 31, 33, 47, 63, 64, 79, 95, 111, 127, 143, 159, 160, 191, 223, 255], False)),
    ("every K, per-base gaps", lengths_batch([15, 47, 79, 111, 143, 175, 207, 255], True)),
    ("lane-aligned + packed (100 = 10 x 10, 150 = 15 x 10)", pkg.Batch.concat([S.fixed_shape(2, 100, 120, 300, 3, 5), S.fixed_shape(1, 150, 170, 40, 2, 6)])),
    ("S4-like: all rescued", S.s4(1, n_reads=8, n_haps=3, hap_lo=200, hap_hi=260)),
    ("long reads (256..600)", lengths_batch([256, 300, 600, 40], False)),
]
with pkg.PairHMMEngine(devices=[0]) as eng, pkg.PairHMMEngine(devices=[0], exact_fp32=True) as ex, \
        pkg.PairHMMEngine(devices=[0], fp64_first=2) as f64, pkg.PairHMMEngine(devices=[0, 0]) as two:
    for name, b in batches:
        check(eng.compute(b), b, name)
        check(ex.compute(b), b, name + " [exact]")
        check(f64.compute(b), b, name + " [FP64 first]")
        check(two.compute(b), b, name + " [two workers]")
    # const gaps with i != d (MODE 1)
    b = S.random_small(9, n_regions=3, max_reads=8, max_haps=3, general_gaps=False)
    b1 = pkg.Batch(b.region_read_beg, b.region_hap_beg, b.read_off, b.read_bases, b.read_q, b.hap_off, b.hap_bases,
                   gap_open_i=ord("I"), gap_open_d=ord("F"), gap_cont_c=ord("-"))
    check(eng.compute(b1), b1, "constant gaps, i != d (MODE 1)")
    # device genotype reduction
    b = S.random_small(11, n_regions=6, max_reads=10, max_haps=5, general_gaps=False)
    per_site = [(g, 3, rng.integers(0, 3, int(b.haps_per_region[g])).astype(np.uint8), (rng.random(int(b.reads_per_region[g])) > 0.3).astype(np.uint8))
                for g in range(b.n_regions) for _ in range(2)]
    gl = eng.compute_gl(b, pkg.Sites(b, per_site), want_matrix=True)
    DUMP["gl"] = np.concatenate([gl.gl.view(np.uint8), gl.capped.view(np.uint8), gl.read_keep, gl.site_n_reads.view(np.uint8)])
    guards = [e.debug_check() for e in (eng, ex, f64, two)]
    print("guard bytes overwritten:", guards, flush=True)
    assert guards == [0, 0, 0, 0]
    assert not np.isnan(gl.gl).any()          # (-inf is legitimate: an allele no haplotype carries)
    print(f"ok  genotype reduction: {len(per_site)} sites", flush=True)
pairs = sw_cases(5, 12)
out, ms = pkg.sw_align(pairs)
DUMP["sw"] = np.frombuffer(repr(out).encode(), np.uint8)
print(f"ok  smith-waterman: {len(out)} alignments", flush=True)
if "--dump" in sys.argv:
    np.savez(sys.argv[sys.argv.index("--dump") + 1], **{k.replace("/", "_"): v for k, v in DUMP.items()})
print("SANITIZE TARGET DONE")
