"""CPU suite, part 1: the oracle (oracle/pairhmm_oracle.c) is pinned to the reference.

Pins, strongest first:
  * tests/golden/*.json -- outputs of the reference's own code (tests/golden/make_golden.py),
    including every known-answer vector of SURVEY.md Appendix A;
  * the compiled reference itself (oracle/_ref), live, when it is present in this checkout.
Bar: bit-exact (float and double), as the reference is deterministic scalar/AVX IEEE arithmetic.
"""
import struct

import numpy as np
import pytest

f32bits = lambda x: struct.unpack("<I", struct.pack("<f", x))[0]
f64bits = lambda x: struct.unpack("<Q", struct.pack("<d", x))[0]


def test_tables_bitwise_oracle_vs_product(pkg, oracle):
    t = pkg.host_tables()
    for name in ("ph2pr_f32", "ph2pr_f64"):
        assert t[name].tobytes() == oracle.table(name).tobytes(), name
    for name in ("mm_f32", "mm_f64"):       # product keeps the reachable prefix (qualities <= 127)
        assert t[name].tobytes() == oracle.table(name)[: len(t[name])].tobytes(), name
    assert len(t["mm_f32"]) == 128 * 129 // 2


def test_tables_bitwise_oracle_vs_reference(oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built in this checkout")
    for name in ("ph2pr_f32", "ph2pr_f64", "mm_f32", "mm_f64"):
        assert oracle.table(name).tobytes() == ref.table(name).tobytes(), name
    assert oracle.log10_init() == ref.log10_init()


def test_table_known_values(oracle, golden):
    # SURVEY.md Appendix A constants
    ph, mm = oracle.table("ph2pr_f32"), oracle.table("mm_f32")
    assert ph[43] == np.float32(5.01187023e-05) and ph[73] == np.float32(5.01187003e-08)
    assert mm[(73 * 74) // 2 + 73] == np.float32(0.999999881)
    assert np.float32(1.0) - ph[43] == np.float32(0.999949872)
    l32, l64 = oracle.log10_init()
    assert l32 == golden["kat_appendix_a"]["log10_init_f32"] == float(np.float32(36.1236000))
    assert l64 == golden["kat_appendix_a"]["log10_init_f64"] == 307.050595577260822


def test_kat_appendix_a1_kernel_level(oracle, golden):
    for k in golden["kat_appendix_a"]["a1"]:
        n = len(k["read"])
        args = (k["read"].encode(), k["qual"].encode(), b"I" * n, b"I" * n, b"+" * n, k["hap"].encode())
        f = oracle.forward(*args, "f32"); d = oracle.forward(*args, "f64")
        assert f32bits(f) == k["f32_bits"], k["name"]
        assert f64bits(d) == k["f64_bits"], k["name"]
        assert int(np.float32(f) < np.float32(1e-28)) == k["rescue"], k["name"]


def test_golden_random_pairs(pkg, oracle, golden):
    for g in golden["ref_random_pairs"]["batches"]:
        b = pkg.synth.random_small(g["seed"], **g["kw"])
        assert b.n_pairs == g["n_pairs"]
        out = oracle.batch(b)
        assert out["raw32"].view(np.uint32).tolist() == g["raw32_bits"]
        assert out["raw64"].view(np.uint64).tolist() == g["raw64_bits"]
        assert out["rescued"].tolist() == g["rescued"]
        assert out["log10"].view(np.uint64).tolist() == g["log10_bits"]


def test_oracle_vs_reference_live(pkg, oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built in this checkout")
    for seed in range(40, 46):
        b = pkg.synth.random_small(seed, n_regions=4, max_read_len=180, max_hap_len=420, lower_frac=0.02)
        a, r = oracle.batch(b), ref.batch(b)
        for key, ty in (("raw32", np.uint32), ("raw64", np.uint64), ("log10", np.uint64)):
            assert np.array_equal(a[key].view(ty), r[key].view(ty)), (seed, key)
        assert np.array_equal(a["rescued"], r["rescued"])
    b = pkg.synth.s4(n_regions=1, n_reads=6, n_haps=2)       # long, every pair rescued
    a, r = oracle.batch(b), ref.batch(b, threads=2)
    assert a["rescued"].all() and np.array_equal(a["log10"].view(np.uint64), r["log10"].view(np.uint64))


def test_threads_do_not_change_results(pkg, oracle):
    b = pkg.synth.s2(2)
    a, c = oracle.batch(b, threads=1), oracle.batch(b, threads=4)
    assert np.array_equal(a["log10"].view(np.uint64), c["log10"].view(np.uint64))


def _region_batch(pkg, reg):
    return pkg.Batch.from_regions([([r.encode() for r in reg["reads"]], [q.encode() for q in reg["quals"]],
                                    [h.encode() for h in reg["haps"]])])


def test_cap_and_filter_vs_reference_call_surface(pkg, oracle, golden):
    """oracle dispatch + oracle cap/filter == hc::IntelPairHMM::compute_likelihoods (A.2 + 6 regions)."""
    import ctypes as C
    a2 = golden["kat_appendix_a"]["a2"]
    regions = [dict(a2, lik_bits=np.array(a2["lik"]).reshape(-1).view(np.uint64).tolist())] + golden["ref_region_filter"]["regions"]
    for reg in regions:
        b = _region_batch(pkg, reg)
        lik = oracle.batch(b)["log10"].reshape(b.n_reads, b.n_haps).copy()
        keep = np.zeros(b.n_reads, np.uint8)
        rl = np.diff(b.read_off).astype(np.int32)
        n = oracle.lib.oracle_normalize_filter(lik.reshape(-1), b.n_reads, b.n_haps, rl, keep)
        assert keep.tolist() == reg["keep"] and n == sum(reg["keep"])
        assert lik[keep.astype(bool)].reshape(-1).view(np.uint64).tolist() == reg["lik_bits"]
        # the product's host-side cap/filter (phmm_normalize_filter) must agree bit for bit
        lik2 = oracle.batch(b)["log10"].reshape(b.n_reads, b.n_haps).copy()
        keep2 = pkg.normalize_filter(lik2, rl)
        assert keep2.tolist() == reg["keep"]


def test_a2_values(golden):
    lik = golden["kat_appendix_a"]["a2"]["lik"]
    want = [[-1.602085114, -6.102085114], [-6.102123260, -1.602123260]]
    assert np.allclose(lik, want, atol=5e-9)
