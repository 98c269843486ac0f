"""per-read allele marginalisation + diploid genotype likelihoods.

CPU : oracle/genotype_oracle.c is bit-identical to the reference's own Genetyper (fixture written by
      tests/golden/make_gl_golden.py from the compiled reference; live when oracle/_ref is present).
GPU : the engine's device-side reduction (phmm_submit_gl / phmm_wait_gl) is bit-identical to the oracle fed with
      the engine's own capped matrix -- and so to the reference -- on every workload shape."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_gl_golden import gl_case  # noqa: E402


def test_oracle_matches_reference_fixture(oracle, golden):
    for c in golden["ref_gl"]["cases"]:
        out, n = oracle.genotype_likelihoods(*gl_case(c["seed"]))
        assert n == c["n_used"], c["seed"]
        assert out.view(np.uint64).tolist() == c["gl_bits"], c["seed"]


def test_oracle_matches_reference_live(oracle, ref):
    if ref is None or not hasattr(ref, "genotype_likelihoods"):
        pytest.skip("oracle/_ref not built")
    for seed in range(200, 260):
        case = gl_case(seed)
        a, na = oracle.genotype_likelihoods(*case)
        b, nb = ref.genotype_likelihoods(*case)
        assert na == nb and np.array_equal(a.view(np.uint64), b.view(np.uint64)), seed


def test_product_jacobian_table_is_the_reference_s(pkg, oracle, golden):
    """The table the device indexes == the one compiled into the reference's binary (hash in the fixture) == the
    oracle's independent binary128 computation == correctly rounded values by mpmath (sampled)."""
    import ctypes as C
    import hashlib
    tab = pkg.jacobian_table()
    assert tab.size == 80001
    assert hashlib.sha256(tab.tobytes()).hexdigest() == golden["ref_gl"]["jacobian_sha256"]
    n = C.c_int()
    otab = np.ctypeslib.as_array(oracle.lib.oracle_jacobian_table(C.byref(n)), (n.value,))
    assert np.array_equal(tab.view(np.uint64), otab.view(np.uint64))
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    for k in list(range(0, 80001, 997)) + [3, 6, 9, 80000]:
        x = -0.0001 * k
        p = float(mp.power(10, mp.mpf(x)))                      # correctly rounded 10^x of the DOUBLE x
        want = float(mp.log10(mp.mpf(1.0 + p)))                 # correctly rounded log10 of the DOUBLE 1 + p
        assert tab[k] == want, k


def random_sites(pkg, b, seed, with_overlap=True, max_sites=4):
    rng = np.random.default_rng(seed)
    per_site = []
    nh, nr = b.haps_per_region, b.reads_per_region
    for g in range(b.n_regions):
        for _ in range(int(rng.integers(0, max_sites + 1))):
            A = int(rng.integers(1, 8))
            ha = rng.integers(0, A, int(nh[g])).astype(np.uint8)
            ov = (rng.random(int(nr[g])) > 0.25).astype(np.uint8) if with_overlap else None
            per_site.append((g, A, ha, ov))
    return pkg.Sites(b, per_site), per_site


def expected_gl(pkg, oracle, b, per_site, log10):
    """Host path: the engine's matrix -> cap / filter (intel_pairhmm.hpp:24-46) -> the oracle's reduction."""
    ob = b.region_out_beg
    nh, nr = b.haps_per_region, b.reads_per_region
    rl = np.diff(b.read_off).astype(np.int32)
    capped, keeps = {}, {}
    out, n_used = [], []
    for g, A, ha, ov in per_site:
        if g not in capped:
            m = log10[int(ob[g]):int(ob[g + 1])].reshape(int(nr[g]), int(nh[g])).copy()
            r0 = int(b.region_read_beg[g])
            keeps[g] = pkg.normalize_filter(m, rl[r0:r0 + int(nr[g])]) if m.size else np.ones(int(nr[g]), np.uint8)
            capped[g] = m
        gl, n = oracle.genotype_likelihoods(capped[g], keeps[g], ov, A, ha) if capped[g].size else (np.zeros(A * (A + 1) // 2), 0)
        out.append(gl); n_used.append(n)
    return out, n_used, capped, keeps


@pytest.mark.gpu
@pytest.mark.parametrize("name,make,overlap", [
    ("S3", lambda S: S.s3(2), True),
    ("S5 ragged windows", lambda S: S.s5_batch(40, seed=21), True),
    ("S4: every pair FP64-rescued", lambda S: S.s4(2, n_reads=24, n_haps=5), False),
    ("hopeless reads (filtered) and tiny regions", lambda S: S.random_small(77, n_regions=25, max_reads=30, max_haps=9, max_read_len=120,
                                                                            max_hap_len=200, general_gaps=False), True),
    ("per-base gap penalties", lambda S: S.random_small(78, n_regions=10, max_reads=12, max_haps=6), True),
])
def test_device_genotype_likelihoods_are_bit_identical(pkg, engine, oracle, name, make, overlap):
    b = make(pkg.synth)
    sites, per_site = random_sites(pkg, b, 5, with_overlap=overlap)
    host = engine.compute(b)
    want, want_n, capped, keeps = expected_gl(pkg, oracle, b, per_site, host.log10)
    for devs in ([0], [0, 0, 0]):                                     # one worker, and the sharded scheduler
        with pkg.PairHMMEngine(devices=devs, pipeline_depth=2) as eng:
            got = eng.compute_gl(b, sites, want_matrix=True)
            assert got.stats["n_pairs"] == b.n_pairs and got.stats["n_rescued"] == host.stats["n_rescued"]
            for k, (g, A, ha, ov) in enumerate(per_site):
                assert got.site(k).view(np.uint64).tolist() == want[k].view(np.uint64).tolist(), (name, devs, k, g)
                assert got.site_n_reads[k] == want_n[k]
            ob = b.region_out_beg
            for g, m in capped.items():                               # the capped matrix and the filter, too
                assert np.array_equal(got.capped[int(ob[g]):int(ob[g + 1])].view(np.uint64), m.reshape(-1).view(np.uint64)), (name, g)
                r0 = int(b.region_read_beg[g])
                assert np.array_equal(got.read_keep[r0:r0 + len(keeps[g])], keeps[g])
            # without the matrix the download is the per-site vectors, the keep flags and 16 bytes of counters
            lean = eng.compute_gl(b, sites)
            assert np.array_equal(lean.gl.view(np.uint64), got.gl.view(np.uint64))
            assert lean.stats["d2h_bytes"] < 16 * len(devs) + 8 * lean.gl.size + 4 * sites.n_sites + b.n_reads + 1024 * len(devs)
            # two tickets in flight
            t1, t2 = eng.submit_gl(b, sites), eng.submit_gl(b, sites)
            for t in (t1, t2):
                assert np.array_equal(eng.wait_gl(t, pkg.GlResult(b, sites)).gl.view(np.uint64), got.gl.view(np.uint64))


@pytest.mark.gpu
def test_gl_ticket_errors(pkg, engine):
    b = pkg.synth.s3(1)
    sites, _ = random_sites(pkg, b, 1)
    t = engine.submit_gl(b, sites)
    with pytest.raises(pkg.PhmmError) as ei:
        engine.wait(t)                                                # a GL ticket is waited with wait_gl
    assert ei.value.code == 6
    engine.wait_gl(t, pkg.GlResult(b, sites))
    bad = pkg.Sites(b, [(0, 2, np.full(16, 3, np.uint8), None)])      # allele 3 of a 2-allele site
    with pytest.raises(pkg.PhmmError) as ei:
        engine.compute_gl(b, bad)
    assert ei.value.code == 1
