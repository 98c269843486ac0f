"""ctypes loaders for the CPU checkers under oracle/ (test infrastructure, never the product).

`load_oracle()`  -> oracle/liboracle.so          (plain-C restatement, oracle/pairhmm_oracle.c)
`load_ref()`     -> oracle/_ref/libref_pairhmm.so (the reference's own code, compiled in place from
                    /root/reference by oracle/Makefile; None if that build is absent)
Both expose the same batch entry point (`*_batch`) over the include/phmm.h batch layout.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")

_PAIR_ARGS = [_u8p, _u8p, _u8p, _u8p, _u8p, C.c_int, _u8p, C.c_int]
_BATCH_ARGS = [C.c_int, _i32p, _i32p, _i32p, _u8p, _u8p, _u8p, _u8p, _u8p, _i32p, _u8p,
               _f64p, _f32p, _f64p, _u8p, C.c_int]


def build_oracle():
    """(Re)build oracle/liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


def _bind(lib, prefix):
    getattr(lib, prefix + "_init").restype = None
    f32 = getattr(lib, prefix + "_forward_f32"); f32.argtypes = _PAIR_ARGS; f32.restype = C.c_float
    f64 = getattr(lib, prefix + "_forward_f64"); f64.argtypes = _PAIR_ARGS; f64.restype = C.c_double
    b = getattr(lib, prefix + "_batch"); b.argtypes = _BATCH_ARGS; b.restype = C.c_int
    for name, ty in (("ph2pr_f32", C.c_float), ("ph2pr_f64", C.c_double),
                     ("mm_f32", C.c_float), ("mm_f64", C.c_double)):
        getattr(lib, f"{prefix}_{name}").restype = C.POINTER(ty)
    getattr(lib, prefix + "_log10_init_f32").restype = C.c_float
    getattr(lib, prefix + "_log10_init_f64").restype = C.c_double
    getattr(lib, prefix + "_init")()
    return lib


class Checker:
    """Uniform python face of liboracle.so / libref_pairhmm.so."""

    MM_SIZE = (255 * 256) // 2

    def __init__(self, lib, prefix):
        self.lib, self.prefix = _bind(lib, prefix), prefix

    def _fn(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def table(self, name):
        n = 128 if name.startswith("ph2pr") else self.MM_SIZE
        ptr = self._fn(name)()
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    def log10_init(self):
        return float(self._fn("log10_init_f32")()), float(self._fn("log10_init_f64")())

    @staticmethod
    def _b(x):
        return np.ascontiguousarray(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else x,
                                    dtype=np.uint8)

    def forward(self, rs, q, i, d, c, hap, dtype="f32"):
        rs, q, i, d, c, hap = map(self._b, (rs, q, i, d, c, hap))
        return self._fn("forward_" + dtype)(rs, q, i, d, c, len(rs), hap, len(hap))

    def batch(self, batch, threads=1):
        """batch: gatk-haplotypecaller-cpp17_b200.Batch-like object (numpy SoA).  Returns dict."""
        n = int(batch.n_pairs)
        out = np.empty(n, np.float64); raw32 = np.empty(n, np.float32)
        raw64 = np.empty(n, np.float64); resc = np.empty(n, np.uint8)
        rc = self._fn("batch")(batch.n_regions, batch.region_read_beg, batch.region_hap_beg,
                               batch.read_off, batch.read_bases, batch.read_q, batch.read_i,
                               batch.read_d, batch.read_c, batch.hap_off, batch.hap_bases,
                               out, raw32, raw64, resc, threads)
        assert rc == 0
        return {"log10": out, "raw32": raw32, "raw64": raw64, "rescued": resc}


def _genotype(fn_oracle, fn_ref, lik, keep, use, n_alleles, hap_allele):
    """Genotype likelihoods of one site from a capped region matrix [n_reads][n_haps]."""
    lik = np.ascontiguousarray(lik, np.float64)
    n_reads, n_haps = lik.shape
    hap_allele = np.ascontiguousarray(hap_allele, np.uint8)
    out = np.empty(n_alleles * (n_alleles + 1) // 2, np.float64)
    keep = np.ones(n_reads, np.uint8) if keep is None else np.ascontiguousarray(keep, np.uint8)
    use = np.ones(n_reads, np.uint8) if use is None else np.ascontiguousarray(use, np.uint8)
    if fn_oracle is not None:
        n = fn_oracle(lik, n_reads, n_haps, keep, use, n_alleles, hap_allele, out)
    else:       # the reference takes the matrix the genotyper sees (erased rows gone) and the overlapping indices
        rows = np.nonzero(keep)[0]
        sub = np.ascontiguousarray(lik[rows])
        idx = np.ascontiguousarray(np.nonzero(use[rows])[0], np.int32)
        fn_ref(sub, len(rows), n_haps, idx, len(idx), n_alleles, hap_allele, out)
        n = len(idx)
    return out, int(n)


def load_oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    ch = Checker(lib, "oracle")
    lib.oracle_normalize_filter.argtypes = [_f64p, C.c_int, C.c_int, _i32p, _u8p]
    lib.oracle_normalize_filter.restype = C.c_int
    lib.oracle_genotype_likelihoods.argtypes = [_f64p, C.c_int, C.c_int, _u8p, _u8p, C.c_int, _u8p, _f64p]
    lib.oracle_genotype_likelihoods.restype = C.c_int
    lib.oracle_jacobian_table.argtypes = [C.POINTER(C.c_int)]
    lib.oracle_jacobian_table.restype = C.POINTER(C.c_double)
    ch.genotype_likelihoods = lambda lik, keep, use, n_alleles, hap_allele: _genotype(
        lib.oracle_genotype_likelihoods, None, lik, keep, use, n_alleles, hap_allele)
    return ch


def load_ref():
    path = os.path.join(ORACLE_DIR, "_ref", "libref_pairhmm.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    ch = Checker(lib, "ref")
    lib.ref_compute_likelihoods.argtypes = [C.c_int, _i32p, _u8p, _u8p, C.c_int, _i32p, _u8p, _f64p, _u8p]
    lib.ref_compute_likelihoods.restype = C.c_int
    if hasattr(lib, "ref_genotype_likelihoods"):
        lib.ref_genotype_likelihoods.argtypes = [_f64p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, _u8p, _f64p]
        lib.ref_genotype_likelihoods.restype = C.c_int
        ch.genotype_likelihoods = lambda lik, keep, use, n_alleles, hap_allele: _genotype(
            None, lib.ref_genotype_likelihoods, lik, keep, use, n_alleles, hap_allele)
    return ch
