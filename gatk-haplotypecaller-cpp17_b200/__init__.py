"""B200-native PairHMM forward engine -- Python face of libphmm_b200.so (ctypes over include/phmm.h).

Only the read x haplotype likelihood path of avis9ditiu/gatk-haplotypecaller-cpp17 lives here
(reference: src/haplotypecaller/pairhmm/intel_pairhmm.hpp).  `PairHMMEngine.compute_likelihoods`
mirrors `hc::IntelPairHMM::compute_likelihoods` (:48-56): log10 likelihood matrix [reads][haps],
capped at best-4.5 per read, poorly modelled reads erased (:24-46).

There is NO CPU fallback: importing works without a GPU (so the library's symbols can be checked),
but every compute call raises PhmmError unless an sm_100 device and the compiled extension are
present.  The directory name is not a valid module name; load it with `__graft_entry__.load_package()`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHMM_LIB", os.path.join(_HERE, "libphmm_b200.so"))   # PHMM_LIB: A/B builds while tuning
CSRC = os.path.join(_HERE, "csrc")

PHMM_OK = 0
PHMM_ERR_UNSUPPORTED = 5
ERR_NAMES = {1: "INVALID_ARG", 2: "NO_DEVICE", 3: "CUDA", 4: "OOM", 5: "UNSUPPORTED", 6: "BAD_TICKET"}


class PhmmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"phmm error {code} ({ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


def build(verbose=False):
    """Compile libphmm_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode:
        raise RuntimeError("building libphmm_b200.so failed")
    return LIB_PATH


class _Options(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_devices", C.c_int32), ("devices", C.POINTER(C.c_int32)),
                ("pipeline_depth", C.c_int32), ("exact_fp32", C.c_int32), ("host_threads", C.c_int32),
                ("use_double", C.c_int32), ("fp64_first", C.c_int32), ("recurrence", C.c_int32)]


class _Batch(C.Structure):
    _fields_ = [("n_regions", C.c_int32), ("n_reads", C.c_int32), ("n_haps", C.c_int32),
                ("region_read_beg", C.c_void_p), ("region_hap_beg", C.c_void_p), ("read_off", C.c_void_p),
                ("read_bases", C.c_void_p), ("read_q", C.c_void_p), ("read_i", C.c_void_p),
                ("read_d", C.c_void_p), ("read_c", C.c_void_p), ("hap_off", C.c_void_p),
                ("hap_bases", C.c_void_p),
                ("gap_open_i", C.c_uint8), ("gap_open_d", C.c_uint8), ("gap_cont_c", C.c_uint8),
                ("flags", C.c_uint8)]


class Stats(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_cells", C.c_int64), ("n_rescued", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("kernel_launches", C.c_int32),
                ("n_devices_used", C.c_int32), ("kernel_ms", C.c_float), ("total_ms", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class _Result(C.Structure):
    _fields_ = [("log10_lik", C.c_void_p), ("raw32", C.c_void_p), ("raw64", C.c_void_p),
                ("rescued", C.c_void_p), ("stats", Stats)]


class _Sites(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("site_region", C.c_void_p), ("site_n_alleles", C.c_void_p),
                ("hap_allele", C.c_void_p), ("read_overlap", C.c_void_p)]


class _GlResult(C.Structure):
    _fields_ = [("genotype_lik", C.c_void_p), ("site_n_reads", C.c_void_p), ("read_keep", C.c_void_p),
                ("capped_lik", C.c_void_p), ("stats", Stats)]


EXPORTS = ["phmm_create", "phmm_destroy", "phmm_compute", "phmm_submit", "phmm_wait", "phmm_strerror",
           "phmm_last_error", "phmm_abi_version", "phmm_normalize_filter", "phmm_tables",
           "phmm_stage", "phmm_run_staged", "phmm_run_staged_ex", "phmm_run_staged_pipelined",
           "phmm_fetch_staged", "phmm_free_staged", "phmm_plan", "phmm_sw_align", "phmm_host_register",
           "phmm_host_unregister", "phmm_host_alloc", "phmm_host_free", "phmm_submit_gl", "phmm_wait_gl", "phmm_jacobian_table", "phmm_sw_release", "phmm_debug_check", "phmm_validate"]

class _PlanInfo(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("mode", C.c_int32), ("n_jobs", C.c_int32), ("n_long_pairs", C.c_int32),
                ("haps_per_job", C.c_int32), ("hap_chunks", C.c_int32), ("haps_per_job64", C.c_int32),
                ("hap_chunks64", C.c_int32), ("n_pairs", C.c_int64), ("n_cells", C.c_int64), ("n_shapes", C.c_int32),
                ("shape_g", C.c_int32 * 32), ("shape_k", C.c_int32 * 32),
                ("jobs_ragged", C.c_int32 * 32), ("jobs_aligned", C.c_int32 * 32)]


class _SwBatch(C.Structure):
    _fields_ = [("n", C.c_int32), ("ref_off", C.c_void_p), ("ref_bases", C.c_void_p), ("alt_off", C.c_void_p),
                ("alt_bases", C.c_void_p), ("w_match", C.c_int32), ("w_mismatch", C.c_int32), ("w_open", C.c_int32),
                ("w_extend", C.c_int32)]


class _SwResult(C.Structure):
    _fields_ = [("offset", C.c_void_p), ("elem_beg", C.c_void_p), ("cap_elems", C.c_int64), ("ops", C.c_void_p),
                ("lens", C.c_void_p), ("kernel_ms", C.c_float)]


SW_NEW_PARAMETERS = (200, -150, -260, -11)     # IntelSWAligner::NEW_SW_PARAMETERS, the default of align()


def sw_align(pairs, params=SW_NEW_PARAMETERS, device=0, cap_elems=None):
    """phmm_sw_align: Smith-Waterman haplotype -> reference alignment of a batch of (ref, alt) byte strings,
    as hc::IntelSWAligner::align does one at a time.  Returns ([(offset, cigar string)], kernel_ms)."""
    n = len(pairs)
    refs = [np.frombuffer(bytes(r), np.uint8) for r, _ in pairs]
    alts = [np.frombuffer(bytes(a), np.uint8) for _, a in pairs]
    cat = lambda xs: np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0, np.uint8)
    off = lambda xs: np.concatenate([[0], np.cumsum([len(x) for x in xs])]).astype(np.int32)
    ref_off, alt_off, ref_b, alt_b = off(refs), off(alts), cat(refs), cat(alts)
    worst = int(ref_off[-1]) + int(alt_off[-1]) + 2 * n
    cap = min(worst, 32 * n + 4096) if cap_elems is None else cap_elems
    while True:
        offset = np.zeros(n, np.int32); elem_beg = np.zeros(n + 1, np.int64)
        ops = np.zeros(max(cap, 1), np.uint8); lens = np.zeros(max(cap, 1), np.int32)
        b = _SwBatch(n, ref_off.ctypes.data, ref_b.ctypes.data, alt_off.ctypes.data, alt_b.ctypes.data, *params)
        r = _SwResult(offset.ctypes.data, elem_beg.ctypes.data, cap, ops.ctypes.data, lens.ctypes.data, 0.0)
        rc = lib().phmm_sw_align(device, C.byref(b), C.byref(r))
        if rc == PHMM_ERR_UNSUPPORTED and cap_elems is None and cap < worst and \
                max(len(x) for x in refs + alts) <= 1023:
            cap = worst                                   # unusually fragmented CIGARs: retry with the worst case
            continue
        if rc != PHMM_OK:
            raise PhmmError(rc, lib().phmm_strerror(rc).decode())
        break
    ch = ops.tobytes().decode("latin1")
    out = [(int(offset[k]), "".join(f"{lens[e]}{ch[e]}" for e in range(int(elem_beg[k]), int(elem_beg[k + 1])))) for k in range(n)]
    return out, float(r.kernel_ms)


def sw_align_timed(pairs, params=SW_NEW_PARAMETERS, device=0, repeats=3):
    """Best wall time in seconds of phmm_sw_align itself (arrays in, arrays out) over `repeats` calls: the cost of
    the C ABI without this module's string packing / unpacking."""
    import time
    n = len(pairs)
    refs = [np.frombuffer(bytes(r), np.uint8) for r, _ in pairs]
    alts = [np.frombuffer(bytes(a), np.uint8) for _, a in pairs]
    off = lambda xs: np.concatenate([[0], np.cumsum([len(x) for x in xs])]).astype(np.int32)
    ref_off, alt_off = off(refs), off(alts)
    ref_b, alt_b = np.ascontiguousarray(np.concatenate(refs)), np.ascontiguousarray(np.concatenate(alts))
    cap = 32 * n + 4096
    offset = np.zeros(n, np.int32); elem_beg = np.zeros(n + 1, np.int64)
    ops = np.zeros(cap, np.uint8); lens = np.zeros(cap, np.int32)
    b = _SwBatch(n, ref_off.ctypes.data, ref_b.ctypes.data, alt_off.ctypes.data, alt_b.ctypes.data, *params)
    r = _SwResult(offset.ctypes.data, elem_beg.ctypes.data, cap, ops.ctypes.data, lens.ctypes.data, 0.0)
    best = 1e9
    for _ in range(repeats + 1):
        t0 = time.perf_counter()
        rc = lib().phmm_sw_align(device, C.byref(b), C.byref(r))
        dt = time.perf_counter() - t0
        if rc != PHMM_OK:
            raise PhmmError(rc, lib().phmm_strerror(rc).decode())
        best = min(best, dt)
    return best


_lib = None


def lib():
    """Load libphmm_b200.so; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PhmmError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.phmm_create.argtypes = [C.POINTER(_Options), C.POINTER(C.c_void_p)]
        L.phmm_destroy.argtypes = [C.c_void_p]; L.phmm_destroy.restype = None
        L.phmm_compute.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Result)]
        L.phmm_submit.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(C.c_int64)]
        L.phmm_wait.argtypes = [C.c_void_p, C.c_int64, C.POINTER(_Result)]
        L.phmm_strerror.argtypes = [C.c_int]; L.phmm_strerror.restype = C.c_char_p
        L.phmm_last_error.argtypes = [C.c_void_p]; L.phmm_last_error.restype = C.c_char_p
        L.phmm_normalize_filter.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.phmm_tables.argtypes = [C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_float)),
                                  C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)),
                                  C.POINTER(C.c_int32)]
        L.phmm_stage.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(C.c_void_p)]
        L.phmm_run_staged.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.phmm_run_staged_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                         C.POINTER(C.c_int32)]
        L.phmm_run_staged_pipelined.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32,
                                                C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.phmm_plan.argtypes = [C.POINTER(_Batch), C.c_int32, C.c_int32, C.POINTER(_PlanInfo), C.c_void_p, C.c_int64]
        L.phmm_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.phmm_host_unregister.argtypes = [C.c_void_p]
        L.phmm_submit_gl.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Sites), C.POINTER(C.c_int64)]
        L.phmm_wait_gl.argtypes = [C.c_void_p, C.c_int64, C.POINTER(_GlResult)]
        L.phmm_validate.argtypes = [C.POINTER(_Batch), C.POINTER(_Sites)]
        L.phmm_jacobian_table.argtypes = [C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int32)]
        L.phmm_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.phmm_host_free.argtypes = [C.c_void_p]
        L.phmm_sw_align.argtypes = [C.c_int32, C.POINTER(_SwBatch), C.POINTER(_SwResult)]
        L.phmm_sw_release.argtypes = []; L.phmm_sw_release.restype = None
        L.phmm_debug_check.argtypes = [C.c_void_p]; L.phmm_debug_check.restype = C.c_int64
        L.phmm_fetch_staged.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Result)]
        L.phmm_free_staged.argtypes = [C.c_void_p, C.c_void_p]; L.phmm_free_staged.restype = None
        _lib = L
    return _lib


def host_tables():
    """The host-built probability tables the kernels use (native/Context.h semantics)."""
    L = lib()
    pf, mf = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
    pd, md = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    n = C.c_int32()
    L.phmm_tables(C.byref(pf), C.byref(mf), C.byref(pd), C.byref(md), C.byref(n))
    return {"ph2pr_f32": np.ctypeslib.as_array(pf, (128,)).copy(), "mm_f32": np.ctypeslib.as_array(mf, (n.value,)).copy(),
            "ph2pr_f64": np.ctypeslib.as_array(pd, (128,)).copy(), "mm_f64": np.ctypeslib.as_array(md, (n.value,)).copy()}


def validate(batch, sites=None):
    """phmm_validate: the status code phmm_submit / phmm_submit_gl would refuse the arguments with (0 = fine); no device."""
    cb = batch.c_struct()
    cs = sites.c_struct() if sites is not None else None
    return int(lib().phmm_validate(C.byref(cb), C.byref(cs) if cs is not None else None))


def jacobian_table():
    """hc::MathUtils' Jacobian-logarithm table as the library holds it (utils/math_utils.hpp:17-29)."""
    t, n = C.POINTER(C.c_double)(), C.c_int32()
    lib().phmm_jacobian_table(C.byref(t), C.byref(n))
    return np.ctypeslib.as_array(t, (n.value,)).copy()


class Sites:
    """Variant sites of a batch for the device-side genotype reduction (phmm_sites of include/phmm.h).
    per_site: list of (region, n_alleles, hap_allele uint8[n_haps(region)], read_overlap uint8[n_reads(region)] or None),
    regions non-decreasing.  If any site gives no overlap array, every read is taken to overlap every site."""

    def __init__(self, batch, per_site):
        self.n_sites = len(per_site)
        self.site_region = np.ascontiguousarray([p[0] for p in per_site], np.int32)
        self.site_n_alleles = np.ascontiguousarray([p[1] for p in per_site], np.int32)
        cat = lambda xs: np.ascontiguousarray(np.concatenate(xs), np.uint8) if xs else np.zeros(0, np.uint8)
        self.hap_allele = cat([np.asarray(p[2], np.uint8) for p in per_site])
        self.has_overlap = bool(per_site) and all(p[3] is not None for p in per_site)
        self.read_overlap = cat([np.asarray(p[3], np.uint8) for p in per_site]) if self.has_overlap else None
        nh, nr = batch.haps_per_region, batch.reads_per_region
        assert len(self.hap_allele) == int(nh[self.site_region].sum()) if self.n_sites else True
        if self.has_overlap:
            assert len(self.read_overlap) == int(nr[self.site_region].sum())
        self.gl_off = np.concatenate([[0], np.cumsum(self.site_n_alleles.astype(np.int64) * (self.site_n_alleles + 1) // 2)]).astype(np.int64)

    def c_struct(self):
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None
        s = _Sites()
        s.n_sites = self.n_sites
        s.site_region, s.site_n_alleles, s.hap_allele = p(self.site_region), p(self.site_n_alleles), p(self.hap_allele)
        s.read_overlap = p(self.read_overlap) if self.has_overlap else None
        return s


class GlResult:
    def __init__(self, batch, sites, want_matrix=False):
        self.gl = np.zeros(int(sites.gl_off[-1]), np.float64)
        self.gl_off = sites.gl_off
        self.site_n_reads = np.zeros(sites.n_sites, np.int32)
        self.read_keep = np.zeros(batch.n_reads, np.uint8)
        self.capped = np.zeros(batch.n_pairs, np.float64) if want_matrix else None
        self._c = _GlResult()
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        self._c.genotype_lik, self._c.site_n_reads, self._c.read_keep, self._c.capped_lik = p(self.gl), p(self.site_n_reads), p(self.read_keep), p(self.capped)

    def site(self, k):
        return self.gl[int(self.gl_off[k]):int(self.gl_off[k + 1])]

    @property
    def stats(self):
        return self._c.stats.as_dict()


class Batch:
    """SoA batch of active regions (the phmm_batch layout of include/phmm.h) held in numpy arrays."""

    def __init__(self, region_read_beg, region_hap_beg, read_off, read_bases, read_q, hap_off, hap_bases,
                 read_i=None, read_d=None, read_c=None, gap_open_i=ord("I"), gap_open_d=ord("I"), gap_cont_c=ord("+")):
        a32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
        a8 = lambda x: np.ascontiguousarray(x, dtype=np.uint8)
        self.region_read_beg, self.region_hap_beg = a32(region_read_beg), a32(region_hap_beg)
        self.read_off, self.hap_off = a32(read_off), a32(hap_off)
        self.read_bases, self.read_q, self.hap_bases = a8(read_bases), a8(read_q), a8(hap_bases)
        self.explicit_gaps = read_i is not None
        n = len(self.read_bases)
        # the checkers always want per-base arrays; the engine gets NULL + constants when uniform
        self.read_i = a8(read_i) if read_i is not None else np.full(n, gap_open_i, np.uint8)
        self.read_d = a8(read_d) if read_d is not None else np.full(n, gap_open_d, np.uint8)
        self.read_c = a8(read_c) if read_c is not None else np.full(n, gap_cont_c, np.uint8)
        self.gap_open_i, self.gap_open_d, self.gap_cont_c = gap_open_i, gap_open_d, gap_cont_c
        self.n_regions = len(self.region_read_beg) - 1
        self.n_reads = len(self.read_off) - 1
        self.n_haps = len(self.hap_off) - 1

    @property
    def reads_per_region(self):
        return np.diff(self.region_read_beg)

    @property
    def haps_per_region(self):
        return np.diff(self.region_hap_beg)

    @property
    def region_out_beg(self):
        return np.concatenate([[0], np.cumsum(self.reads_per_region.astype(np.int64) * self.haps_per_region)])

    @property
    def n_pairs(self):
        return int(self.region_out_beg[-1])

    @property
    def n_cells(self):
        # bases per region from the offset arrays (regions without reads or haplotypes contribute nothing)
        rsum = np.diff(np.asarray(self.read_off, np.int64)[self.region_read_beg])
        hsum = np.diff(np.asarray(self.hap_off, np.int64)[self.region_hap_beg])
        return int((rsum * hsum).sum())

    @property
    def input_bytes(self):
        per_base = 5 if self.explicit_gaps else 2
        return int(per_base * len(self.read_bases) + len(self.hap_bases))

    def region_cells(self):
        """Cell count (sum of read_len x hap_len over pairs) of every region."""
        rl = np.concatenate([[0], np.cumsum(np.diff(self.read_off).astype(np.int64))])
        hl = np.concatenate([[0], np.cumsum(np.diff(self.hap_off).astype(np.int64))])
        return (rl[self.region_read_beg[1:]] - rl[self.region_read_beg[:-1]]) * \
               (hl[self.region_hap_beg[1:]] - hl[self.region_hap_beg[:-1]])

    def slice_regions(self, g0, g1):
        """Regions [g0, g1) as their own Batch (offsets rebased); outputs keep their relative order."""
        r0, r1 = int(self.region_read_beg[g0]), int(self.region_read_beg[g1])
        h0, h1 = int(self.region_hap_beg[g0]), int(self.region_hap_beg[g1])
        b0, b1 = int(self.read_off[r0]), int(self.read_off[r1])
        c0, c1 = int(self.hap_off[h0]), int(self.hap_off[h1])
        kw = dict(gap_open_i=self.gap_open_i, gap_open_d=self.gap_open_d, gap_cont_c=self.gap_cont_c)
        if self.explicit_gaps:
            kw.update(read_i=self.read_i[b0:b1], read_d=self.read_d[b0:b1], read_c=self.read_c[b0:b1])
        return Batch(self.region_read_beg[g0:g1 + 1] - r0, self.region_hap_beg[g0:g1 + 1] - h0,
                     self.read_off[r0:r1 + 1] - b0, self.read_bases[b0:b1], self.read_q[b0:b1],
                     self.hap_off[h0:h1 + 1] - c0, self.hap_bases[c0:c1], **kw)

    @staticmethod
    def concat(batches):
        """Several batches as one (regions in order)."""
        batches = list(batches)
        if len(batches) == 1:
            return batches[0]
        b0 = batches[0]
        assert all(b.explicit_gaps == b0.explicit_gaps for b in batches)
        def offs(name):
            out, base = [np.zeros(1, np.int64)], 0
            for b in batches:
                a = np.asarray(getattr(b, name), np.int64)
                out.append(a[1:] + base); base += int(a[-1])
            return np.concatenate(out)
        cat = lambda name: np.concatenate([getattr(b, name) for b in batches])
        kw = dict(gap_open_i=b0.gap_open_i, gap_open_d=b0.gap_open_d, gap_cont_c=b0.gap_cont_c)
        if b0.explicit_gaps:
            kw.update(read_i=cat("read_i"), read_d=cat("read_d"), read_c=cat("read_c"))
        return Batch(offs("region_read_beg"), offs("region_hap_beg"), offs("read_off"), cat("read_bases"), cat("read_q"),
                     offs("hap_off"), cat("hap_bases"), **kw)

    def c_struct(self):
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        b = _Batch()
        b.n_regions, b.n_reads, b.n_haps = self.n_regions, self.n_reads, self.n_haps
        b.region_read_beg, b.region_hap_beg = p(self.region_read_beg), p(self.region_hap_beg)
        b.read_off, b.hap_off = p(self.read_off), p(self.hap_off)
        b.read_bases, b.read_q, b.hap_bases = p(self.read_bases), p(self.read_q), p(self.hap_bases)
        if self.explicit_gaps:
            b.read_i, b.read_d, b.read_c = p(self.read_i), p(self.read_d), p(self.read_c)
        else:
            b.read_i = b.read_d = b.read_c = None
        b.gap_open_i, b.gap_open_d, b.gap_cont_c = self.gap_open_i, self.gap_open_d, self.gap_cont_c
        b.flags = 1 if getattr(self, "_pinned", False) else 0       # PHMM_BATCH_PINNED_INPUTS
        return b

    def pin(self):
        """Page-lock the byte arrays (phmm_host_register) and mark the batch PHMM_BATCH_PINNED_INPUTS: submits
        then upload straight from these arrays (they must stay alive and unchanged until the matching wait)."""
        if not getattr(self, "_pinned", False):
            arrs = [self.read_bases, self.read_q, self.hap_bases] + ([self.read_i, self.read_d, self.read_c] if self.explicit_gaps else [])
            self._pinned_arrays = []                     # every successful registration is recorded as it happens
            for a in arrs:
                if a.nbytes:
                    rc = lib().phmm_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes)
                    if rc != PHMM_OK:
                        self._unregister_all()           # roll back: nothing stays page-locked behind a failed pin()
                        raise PhmmError(rc, lib().phmm_strerror(rc).decode())
                    self._pinned_arrays.append(a)
            self._pinned = True
        return self

    def _unregister_all(self):
        for a in getattr(self, "_pinned_arrays", []):
            lib().phmm_host_unregister(a.ctypes.data_as(C.c_void_p))
        self._pinned_arrays = []

    def unpin(self):
        self._unregister_all()
        self._pinned = False

    def __del__(self):                                   # never free memory that is still cudaHostRegister-ed
        try:
            self._unregister_all()
        except Exception:
            pass

    @staticmethod
    def from_regions(regions, **kw):
        """regions: list of (reads, quals, haps[, ins, del, gcp]) with bytes / uint8 arrays."""
        rrb, rhb, roff, hoff = [0], [0], [0], [0]
        rb, rq, hb, ri, rd_, rc = [], [], [], [], [], []
        explicit = any(len(reg) > 3 for reg in regions)
        u8 = lambda x: np.frombuffer(x, np.uint8) if isinstance(x, (bytes, bytearray)) else np.asarray(x, np.uint8)
        for reg in regions:
            reads, quals, haps = reg[0], reg[1], reg[2]
            for k, (s, q) in enumerate(zip(reads, quals)):
                s, q = u8(s), u8(q)
                assert len(s) == len(q)
                rb.append(s); rq.append(q); roff.append(roff[-1] + len(s))
                if explicit:
                    ri.append(u8(reg[3][k])); rd_.append(u8(reg[4][k])); rc.append(u8(reg[5][k]))
            for h in haps:
                h = u8(h); hb.append(h); hoff.append(hoff[-1] + len(h))
            rrb.append(len(roff) - 1); rhb.append(len(hoff) - 1)
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
        if explicit:
            kw.update(read_i=cat(ri), read_d=cat(rd_), read_c=cat(rc))
        return Batch(rrb, rhb, roff, cat(rb), cat(rq), hoff, cat(hb), **kw)


def plan(batch, sm_count=148, host_threads=1):
    """phmm_plan: the planner alone (no device).  Returns (info dict, jobs int32 array [n_jobs, 10])."""
    info = _PlanInfo()
    info.struct_size = C.sizeof(_PlanInfo)
    cb = batch.c_struct()
    rc = lib().phmm_plan(C.byref(cb), sm_count, host_threads, C.byref(info), None, 0)
    if rc != PHMM_OK:
        raise PhmmError(rc, lib().phmm_strerror(rc).decode())
    jobs = np.zeros((info.n_jobs, 10), np.int32)
    if info.n_jobs:
        rc = lib().phmm_plan(C.byref(cb), sm_count, host_threads, C.byref(info), jobs.ctypes.data, info.n_jobs)
        if rc != PHMM_OK:
            raise PhmmError(rc, lib().phmm_strerror(rc).decode())
    n = info.n_shapes
    d = {k: getattr(info, k) for k in ("mode", "n_jobs", "n_long_pairs", "haps_per_job", "hap_chunks", "haps_per_job64",
                                      "hap_chunks64", "n_pairs", "n_cells", "n_shapes")}
    d["shapes"] = [(info.shape_g[i], info.shape_k[i]) for i in range(n)]
    d["jobs_ragged"] = [info.jobs_ragged[i] for i in range(n)]
    d["jobs_aligned"] = [info.jobs_aligned[i] for i in range(n)]
    return d, jobs


class Result:
    def __init__(self, n_pairs, want_raw=True):
        self.log10 = np.empty(n_pairs, np.float64)
        self.raw32 = np.empty(n_pairs, np.float32) if want_raw else None
        self.raw64 = np.empty(n_pairs, np.float64) if want_raw else None
        self.rescued = np.empty(n_pairs, np.uint8) if want_raw else None
        self._c = _Result()
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        self._c.log10_lik, self._c.raw32, self._c.raw64, self._c.rescued = p(self.log10), p(self.raw32), p(self.raw64), p(self.rescued)

    @property
    def stats(self):
        return self._c.stats.as_dict()


class PairHMMEngine:
    """Owns one phmm_engine (streams, memory pool, worker per device)."""

    def __init__(self, devices=None, pipeline_depth=2, exact_fp32=False, host_threads=1, use_double=False, fp64_first=0, recurrence=0):
        self._L = lib()
        opt = _Options()
        opt.struct_size = C.sizeof(_Options)
        devices = [0] if devices is None else list(devices)
        self._dev_arr = (C.c_int32 * len(devices))(*devices)
        opt.n_devices = len(devices)
        opt.devices = C.cast(self._dev_arr, C.POINTER(C.c_int32))
        opt.pipeline_depth, opt.exact_fp32, opt.host_threads = pipeline_depth, int(exact_fp32), host_threads
        opt.fp64_first = int(fp64_first)                # 0 auto, 1 never, 2 always (include/phmm.h)
        opt.recurrence = int(recurrence)                # 0 scaled recurrence where applicable, 1 reference operation order
        opt.use_double = int(use_double)                # the reference's g_use_double (intel_pairhmm.hpp:58,71,135)
        self._h = C.c_void_p()
        rc = self._L.phmm_create(C.byref(opt), C.byref(self._h))
        if rc != PHMM_OK:
            raise PhmmError(rc, self._L.phmm_strerror(rc).decode())
        self._inflight = {}

    def _check(self, rc):
        if rc != PHMM_OK:
            raise PhmmError(rc, self._L.phmm_strerror(rc).decode() + ": " + self._L.phmm_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.phmm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def debug_check(self):
        """Guard bytes overwritten so far (PHMM_DEBUG_GUARD=1), see include/phmm.h."""
        return int(self._L.phmm_debug_check(self._h))

    # -- phmm_compute / phmm_submit / phmm_wait
    def compute(self, batch, want_raw=True):
        res = Result(batch.n_pairs, want_raw)
        cb = batch.c_struct()
        self._check(self._L.phmm_compute(self._h, C.byref(cb), C.byref(res._c)))
        return res

    def submit(self, batch):
        t = C.c_int64()
        cb = batch.c_struct()
        self._check(self._L.phmm_submit(self._h, C.byref(cb), C.byref(t)))
        self._inflight[t.value] = batch.n_pairs
        return t.value

    def wait(self, ticket, want_raw=False, result=None):
        n = self._inflight.pop(ticket, 0)
        res = result if result is not None else Result(n, want_raw)
        self._check(self._L.phmm_wait(self._h, ticket, C.byref(res._c)))
        return res

    # -- device-side genotype reduction: phmm_submit_gl / phmm_wait_gl
    def submit_gl(self, batch, sites):
        t = C.c_int64()
        cb, cs = batch.c_struct(), sites.c_struct()
        self._check(self._L.phmm_submit_gl(self._h, C.byref(cb), C.byref(cs), C.byref(t)))
        return t.value

    def wait_gl(self, ticket, result):
        self._check(self._L.phmm_wait_gl(self._h, ticket, C.byref(result._c)))
        return result

    def compute_gl(self, batch, sites, want_matrix=False):
        return self.wait_gl(self.submit_gl(batch, sites), GlResult(batch, sites, want_matrix))

    # -- device-resident form
    def stage(self, batch):
        st = C.c_void_p()
        cb = batch.c_struct()
        self._check(self._L.phmm_stage(self._h, C.byref(cb), C.byref(st)))
        return st

    def run_staged(self, st, iters=1):
        ms, n = C.c_float(), C.c_int32()
        self._check(self._L.phmm_run_staged(self._h, st, iters, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def run_staged_ex(self, st, iters=1):
        """(ms per pass, ms of the FP32 forward launch alone or -1, launches per pass)"""
        ms, ms32, n = C.c_float(), C.c_float(), C.c_int32()
        self._check(self._L.phmm_run_staged_ex(self._h, st, iters, C.byref(ms), C.byref(ms32), C.byref(n)))
        return ms.value, ms32.value, n.value

    def run_staged_pipelined(self, staged, steps):
        """`steps` passes, step i on staged[i % n], each batch on its own stream; (total device ms, launches)"""
        arr = (C.c_void_p * len(staged))(*[s.value for s in staged])
        ms, n = C.c_float(), C.c_int32()
        self._check(self._L.phmm_run_staged_pipelined(self._h, arr, len(staged), steps, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def fetch_staged(self, st, n_pairs, want_raw=True):
        res = Result(n_pairs, want_raw)
        self._check(self._L.phmm_fetch_staged(self._h, st, C.byref(res._c)))
        return res

    def free_staged(self, st):
        self._L.phmm_free_staged(self._h, st)

    # -- the reference's call surface: hc::IntelPairHMM::compute_likelihoods (intel_pairhmm.hpp:48-56)
    def compute_likelihoods(self, haplotypes, reads, quals):
        """haplotypes: list of bytes; reads/quals: lists of bytes (SEQ / QUAL, raw Phred+33).
        Returns (matrix [kept][haps] float64, kept_indices).  Reads failing the poorly-modelled
        filter are dropped, as the reference erases them from its `reads` vector (:40-45)."""
        batch = Batch.from_regions([(reads, quals, haplotypes)])
        res = self.compute(batch, want_raw=False)
        lik = res.log10.reshape(len(reads), len(haplotypes)).copy()
        keep = normalize_filter(lik, np.array([len(r) for r in reads], np.int32))
        idx = np.nonzero(keep)[0]
        return lik[idx], idx


def shard_bounds(cells, world):
    """Contiguous split of regions over `world` ranks balanced by cell count (no exchange between
    ranks: every pair is independent, intel_pairhmm.hpp:131-147).  Same rule as the engine's own
    multi-device split (phmm_engine.cu: split_regions).  Returns world+1 region boundaries."""
    pre = np.concatenate([[0], np.cumsum(np.asarray(cells, np.int64))])
    total, n = int(pre[-1]), len(cells)
    cut, g = [0], 0
    for d in range(1, world):
        want = total * d // world
        while g < n and pre[g + 1] <= want:
            g += 1
        if g < n and (want - pre[g]) > (pre[g + 1] - want):
            g += 1
        cut.append(max(cut[-1], g))
    cut.append(n)
    return cut


def shard_regions(batch, rank, world):
    """The sub-batch rank `rank` of `world` computes, and the offset of its outputs in the whole batch."""
    cut = shard_bounds(batch.region_cells(), world)
    g0, g1 = cut[rank], cut[rank + 1]
    return batch.slice_regions(g0, g1), int(batch.region_out_beg[g0])


def normalize_filter(lik, read_len):
    """In-place cap at best-4.5 and poorly-modelled-read flags (intel_pairhmm.hpp:24-46), host side."""
    assert lik.dtype == np.float64 and lik.flags["C_CONTIGUOUS"], "in-place: needs a contiguous float64 matrix"
    n_reads, n_haps = lik.shape
    keep = np.zeros(n_reads, np.uint8)
    rl = np.ascontiguousarray(read_len, np.int32)
    lib().phmm_normalize_filter(lik.ctypes.data_as(C.c_void_p), n_reads, n_haps,
                                rl.ctypes.data_as(C.c_void_p), keep.ctypes.data_as(C.c_void_p))
    return keep


from . import synth  # noqa: E402,F401
