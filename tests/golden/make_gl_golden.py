"""Writes tests/golden/ref_gl.json: genotype likelihoods of random sites computed by the REFERENCE's own
Genetyper::marginal_likelihoods + calculate_genotype_likelihoods (genotyper/genotyper.hpp:245-327), through
oracle/_ref/libref_pairhmm.so (compiled from /root/reference by oracle/Makefile).  Run in the build container:
    python tests/golden/make_gl_golden.py
The cases are regenerated from their seeds by tests/test_genotype.py (gl_case below), only the outputs are stored."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gl_case(seed):
    """(lik [n_reads][n_haps], keep, use, n_alleles, hap_allele): log10 likelihood rows as the engine produces them
    (capped at best - 4.5), with the corner cases of the reduction: alleles no haplotype carries (lowest()),
    -inf entries, allele differences beyond the Jacobian table's 8.0, erased and non-overlapping reads."""
    rng = np.random.default_rng(seed)
    n_reads = int(rng.integers(1, 60))
    n_haps = int(rng.integers(1, 17))
    n_alleles = int(rng.integers(1, 8))
    lik = -rng.random((n_reads, n_haps)) * rng.choice([1.0, 10.0, 300.0]) - 1.6
    lik = np.float32(lik).astype(np.float64) if seed % 3 else lik         # float-valued doubles, as on the FP32 path
    best = lik.max(axis=1, keepdims=True)
    if seed % 4:
        lik = np.maximum(lik, best - 4.5)
    if seed % 7 == 0:
        lik[rng.integers(0, n_reads), :] = -np.inf
    hap_allele = rng.integers(0, n_alleles, n_haps).astype(np.uint8)
    if seed % 5 == 0 and n_alleles > 1:
        hap_allele[hap_allele == n_alleles - 1] = 0                        # an allele without haplotypes
    keep = (rng.random(n_reads) > 0.1).astype(np.uint8)
    use = (rng.random(n_reads) > 0.3).astype(np.uint8)
    if seed % 11 == 0:
        use[:] = 0                                                         # no read overlaps the site
    return lik, keep, use, n_alleles, hap_allele


if __name__ == "__main__":
    from _oracle import load_ref
    ref = load_ref()
    assert ref is not None and hasattr(ref, "genotype_likelihoods"), "build oracle/_ref first (needs /root/reference)"
    cases = []
    for seed in range(1, 121):
        out, n = ref.genotype_likelihoods(*gl_case(seed))
        cases.append({"seed": seed, "n_used": n, "gl_bits": out.view(np.uint64).tolist()})
    import ctypes as C
    import hashlib
    ref.lib.ref_jacobian_table.restype = C.POINTER(C.c_double)
    ref.lib.ref_jacobian_table.argtypes = [C.POINTER(C.c_int)]
    n = C.c_int()
    tab = np.ctypeslib.as_array(ref.lib.ref_jacobian_table(C.byref(n)), (n.value,)).copy()
    json.dump({"jacobian_sha256": hashlib.sha256(tab.tobytes()).hexdigest(), "jacobian_size": int(n.value),
               "jacobian_note": "hc::MathUtils::JacobianLogTable::cache as compiled into the reference's binary (GCC folds it at compile time)",
               "source": "hc::Genetyper::marginal_likelihoods + calculate_genotype_likelihoods, compiled from /root/reference",
               "cases": cases}, open(os.path.join(ROOT, "tests", "golden", "ref_gl.json"), "w"))
    print(f"wrote {len(cases)} cases")
