"""GPU suite (-m gpu): the CUDA path through the C ABI (phmm_compute / submit / wait / staged) against
the oracle and the committed golden fixtures.

Bars (BASELINE.json north_star): |log10 - oracle| <= 1e-4 on the FP32 path, <= 1e-9 on FP64-rescued
pairs, identical rescue decisions; and with exact_fp32=1 the raw FP32 forward sums are BIT-IDENTICAL
to the reference (and so is every rescue decision by construction).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL32, TOL64 = 1e-4, 1e-9


def _maxerr(a, b):
    if not len(a):
        return 0.0
    d = np.where(a == b, 0.0, np.abs(a - b))       # equal infinities count as zero
    return float(np.nan_to_num(d, nan=np.inf).max())


def check(got, want, exact=False, what=""):
    resc = want["rescued"].astype(bool)
    assert np.array_equal(got.rescued.astype(bool), resc), f"{what}: rescue decisions differ"
    assert _maxerr(got.log10[~resc], want["log10"][~resc]) <= TOL32, what
    assert _maxerr(got.log10[resc], want["log10"][resc]) <= TOL64, what
    if exact:
        assert np.array_equal(got.raw32.view(np.uint32), want["raw32"].view(np.uint32)), f"{what}: raw FP32 bits"
        assert np.array_equal(got.log10[~resc].view(np.uint64), want["log10"][~resc].view(np.uint64)), what
    assert got.stats["kernel_launches"] > 0 and got.stats["n_pairs"] == len(want["log10"])
    assert got.stats["n_rescued"] == int(resc.sum())


def _kat_batch(pkg, kats):
    return pkg.Batch.from_regions([([k["read"].encode()], [k["qual"].encode()], [k["hap"].encode()]) for k in kats])


def test_kat_appendix_a_through_the_c_abi(pkg, engine, exact_engine, golden):
    kats = golden["kat_appendix_a"]["a1"]
    b = _kat_batch(pkg, kats)
    ex = exact_engine.compute(b)
    assert ex.raw32.view(np.uint32).tolist() == [k["f32_bits"] for k in kats]
    assert ex.rescued.tolist() == [k["rescue"] for k in kats]
    fast = engine.compute(b)
    assert fast.rescued.tolist() == [k["rescue"] for k in kats]
    l32, l64 = golden["kat_appendix_a"]["log10_init_f32"], golden["kat_appendix_a"]["log10_init_f64"]
    for k, v, e in zip(kats, fast.log10, ex.log10):
        if k["rescue"]:
            want = np.log10(np.array([k["f64_bits"]], np.uint64).view(np.float64)[0]) - l64
            assert abs(v - want) <= TOL64 and abs(e - want) <= TOL64, k["name"]
        else:
            f = np.array([k["f32_bits"]], np.uint32).view(np.float32)[0]
            want = float(np.float32(np.log10(f, dtype=np.float32)) - np.float32(l32))
            assert abs(v - want) <= TOL32 and abs(e - want) <= TOL32, k["name"]


def test_golden_reference_pairs(pkg, engine, exact_engine, golden):
    for g in golden["ref_random_pairs"]["batches"]:
        b = pkg.synth.random_small(g["seed"], **g["kw"])
        want = {"raw32": np.array(g["raw32_bits"], np.uint32).view(np.float32),
                "log10": np.array(g["log10_bits"], np.uint64).view(np.float64),
                "rescued": np.array(g["rescued"], np.uint8)}
        check(engine.compute(b), want, what=f"golden seed {g['seed']} fast")
        check(exact_engine.compute(b), want, exact=True, what=f"golden seed {g['seed']} exact")


@pytest.mark.parametrize("seed,kw", [
    (1, dict(n_regions=6)),                                              # per-base gap penalties, N bases
    (2, dict(n_regions=6, general_gaps=False)),                          # the reference's constant 'I','I','+'
    (3, dict(n_regions=8, max_read_len=255, max_hap_len=600)),           # longest single-pass read, long haps
    (4, dict(n_regions=5, lower_frac=0.1, n_frac=0.1)),                  # lower case -> 'A', many N
    (5, dict(n_regions=12, max_reads=1, max_haps=1)),                    # 1 x 1 regions (odd counts everywhere)
    (6, dict(n_regions=4, max_read_len=3, max_hap_len=3)),               # tiny R and H, reads longer than haps
    (7, dict(n_regions=3, max_reads=40, max_haps=20, general_gaps=False)),
])
def test_random_ragged_batches(pkg, engine, exact_engine, oracle, seed, kw):
    b = pkg.synth.random_small(seed, **kw)
    want = oracle.batch(b, threads=8)
    check(engine.compute(b), want, what=f"seed {seed} fast")
    check(exact_engine.compute(b), want, exact=True, what=f"seed {seed} exact")


@pytest.mark.parametrize("general,with_n", [(False, False), (False, True), (True, True)])
def test_every_read_length_1_to_255(pkg, engine, exact_engine, oracle, general, with_n):
    """One read of every length 1..255: every compiled lane-group shape with every number of dummy rows,
    lane-aligned and not, odd read counts per shape (idle lane groups, unpaired reads), haplotypes with and
    without N, and haplotypes shorter than the lane group (several in flight at once).  The long reads
    against the short haplotypes end near the bottom of the FP64 range, where the reference's
    flush-to-zero of DOUBLE denormals shows (precision tier 3 of phmm_kernels.cuh)."""
    rng = np.random.default_rng(77 + general + 2 * with_n)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    regions = []
    for reg in range(2):
        haps = [alpha[rng.integers(0, 4, int(n))] for n in (311, 97, 5)]
        if with_n:
            haps[1] = haps[1].copy(); haps[1][::17] = ord("N")
        lens = np.arange(1 + reg, 256, 2)                  # odd lengths in region 0, even in region 1
        rng.shuffle(lens)
        reads, quals, gi, gd, gc = [], [], [], [], []
        for rl in lens:
            rl = int(rl)
            h = haps[0]
            o = int(rng.integers(0, len(h) - rl + 1)); r = h[o:o + rl].copy()
            m = rng.random(rl) < 0.03; r[m] = alpha[rng.integers(0, 4, int(m.sum()))]
            reads.append(r)
            quals.append((33 + rng.integers(2, 42, rl)).astype(np.uint8))
            gi.append((33 + rng.integers(20, 50, rl)).astype(np.uint8))
            gd.append((33 + rng.integers(20, 50, rl)).astype(np.uint8))
            gc.append((33 + rng.integers(5, 25, rl)).astype(np.uint8))
        regions.append((reads, quals, haps, gi, gd, gc) if general else (reads, quals, haps))
    b = pkg.Batch.from_regions(regions)
    want = oracle.batch(b, threads=16)
    check(engine.compute(b), want, what="all lengths fast")
    check(exact_engine.compute(b), want, exact=True, what="all lengths exact")


@pytest.mark.parametrize("general", [False, True])
def test_reads_longer_than_one_lane_group_pass(pkg, engine, exact_engine, oracle, general):
    """Reads of 256..2048 bases take the one-warp-per-pair kernel (phmm_long.cu); mixed in one region with
    short reads (register-tiled kernels), against haplotypes longer and shorter than the reads, with N.
    Good matches stay FP32, poor ones are redone in FP64 inside the same launch."""
    rng = np.random.default_rng(300 + general)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    regions = []
    for reg in range(2):
        haps = [alpha[rng.integers(0, 4, int(n))] for n in (2300, 700, 40)]
        haps[1] = haps[1].copy(); haps[1][::53] = ord("N")
        haps.append(np.concatenate([haps[0][:1000], haps[0][1003:]]))      # a 3-base deletion of hap 0
        lens = [256, 257, 300, 511, 512, 513, 777, 1024, 1500, 2047, 2048, 100, 255, 31] if reg == 0 else [256, 1025, 64, 2048]
        reads, quals, gi, gd, gc = [], [], [], [], []
        for rl in lens:
            h = haps[0]
            o = int(rng.integers(0, len(h) - rl + 1)); r = h[o:o + rl].copy()
            m = rng.random(rl) < 0.01; r[m] = alpha[rng.integers(0, 4, int(m.sum()))]
            reads.append(r)
            quals.append((33 + rng.integers(20, 42, rl)).astype(np.uint8))
            gi.append((33 + rng.integers(30, 50, rl)).astype(np.uint8))
            gd.append((33 + rng.integers(30, 50, rl)).astype(np.uint8))
            gc.append((33 + rng.integers(5, 25, rl)).astype(np.uint8))
        regions.append((reads, quals, haps, gi, gd, gc) if general else (reads, quals, haps))
    b = pkg.Batch.from_regions(regions)
    want = oracle.batch(b, threads=16)
    assert want["rescued"].any() and not want["rescued"].all()
    check(engine.compute(b), want, what="long reads fast")
    check(exact_engine.compute(b), want, exact=True, what="long reads exact")


def test_constant_but_unequal_gap_penalties(pkg, engine, exact_engine, oracle):
    """Batch-constant (i,d,c) with i != d takes kernel MODE 1; NULL arrays and explicit constant arrays agree."""
    b = pkg.synth.random_small(31, n_regions=4, general_gaps=False, max_read_len=255, max_hap_len=300)
    kw = dict(gap_open_i=ord("I"), gap_open_d=ord("F"), gap_cont_c=ord("-"))
    b1 = pkg.Batch(b.region_read_beg, b.region_hap_beg, b.read_off, b.read_bases, b.read_q, b.hap_off, b.hap_bases, **kw)
    n = len(b.read_bases)
    b2 = pkg.Batch(b.region_read_beg, b.region_hap_beg, b.read_off, b.read_bases, b.read_q, b.hap_off, b.hap_bases,
                   read_i=np.full(n, ord("I"), np.uint8), read_d=np.full(n, ord("F"), np.uint8), read_c=np.full(n, ord("-"), np.uint8))
    want = oracle.batch(b1)
    check(exact_engine.compute(b1), want, exact=True, what="mode 1 NULL arrays")
    check(exact_engine.compute(b2), want, exact=True, what="mode 1 explicit arrays")
    check(engine.compute(b2), want, what="mode 1 fast")


@pytest.mark.parametrize("name,make", [
    ("S2 100x300 64x8", lambda s: s.s2(4)),
    ("S3 150x500 256x16", lambda s: s.s3(2)),
    ("S3 general gaps", lambda s: s.s3(1, general_gaps=True)),
    ("S4 long, all rescued", lambda s: s.s4(2, n_reads=24, n_haps=4)),
    ("S4 general gaps", lambda s: s.s4(1, n_reads=12, n_haps=3, general_gaps=True)),
    ("S5 window stream", lambda s: next(s.s5_stream(6, windows_per_batch=6))),
])
def test_baseline_shapes_vs_oracle(pkg, engine, exact_engine, oracle, name, make):
    b = make(pkg.synth)
    want = oracle.batch(b, threads=16)
    check(engine.compute(b), want, what=name)
    check(exact_engine.compute(b), want, exact=True, what=name + " exact")
    if name.startswith("S4"):
        assert want["rescued"].all()


def test_full_size_s3_properties(pkg, engine, oracle):
    """BASELINE config 3 at full size (64 regions, 1.97e10 cells): size-independent properties plus an
    oracle spot check (the oracle would need minutes for the whole batch)."""
    b = pkg.synth.s3(64)
    got = engine.compute(b)
    assert got.stats["n_cells"] == 64 * 256 * 16 * 150 * 500
    # (1) batch-split invariance: every region computed alone gives the same bits
    for g in (0, 31, 63):
        alone = engine.compute(b.slice_regions(g, g + 1))
        o = int(b.region_out_beg[g])
        assert np.array_equal(alone.log10.view(np.uint64), got.log10[o:o + alone.log10.size].view(np.uint64))
    # (2) read-order invariance inside a region: pairs are independent, whatever lane partner they get
    reg = b.slice_regions(5, 6)
    nr, nh, R = 256, 16, 150
    perm = np.random.default_rng(0).permutation(nr)
    shuf = pkg.Batch(reg.region_read_beg, reg.region_hap_beg, reg.read_off,
                     reg.read_bases.reshape(nr, R)[perm].reshape(-1), reg.read_q.reshape(nr, R)[perm].reshape(-1),
                     reg.hap_off, reg.hap_bases)
    a = engine.compute(reg).log10.reshape(nr, nh)
    s = engine.compute(shuf).log10.reshape(nr, nh)
    assert np.array_equal(a[perm].view(np.uint64), s.view(np.uint64))
    # (3) duplicated haplotype -> duplicated column
    dup = pkg.Batch(reg.region_read_beg, [0, 2], reg.read_off, reg.read_bases, reg.read_q, [0, 500, 1000],
                    np.concatenate([reg.hap_bases[:500], reg.hap_bases[:500]]))
    d = engine.compute(dup).log10.reshape(nr, 2)
    assert np.array_equal(d[:, 0].view(np.uint64), d[:, 1].view(np.uint64))
    assert np.array_equal(d[:, 0].view(np.uint64), a[:, 0].view(np.uint64))
    # (4) oracle spot check on 3 whole regions
    for g in (7, 40):
        sub = b.slice_regions(g, g + 1)
        want = oracle.batch(sub, threads=16)
        o = int(b.region_out_beg[g])
        resc = want["rescued"].astype(bool)
        assert np.array_equal(got.rescued[o:o + sub.n_pairs].astype(bool), resc)
        assert _maxerr(got.log10[o:o + sub.n_pairs][~resc], want["log10"][~resc]) <= TOL32
        assert _maxerr(got.log10[o:o + sub.n_pairs][resc], want["log10"][resc]) <= TOL64


@pytest.mark.parametrize("name,make,spot", [
    ("S2 full size: 256 regions of 64 reads (100) x 8 haplotypes (300)", lambda S: S.s2(256), (0, 100, 255)),
    ("S4 full size: 16 regions of 128 reads (150-250, low-quality tails) x 16 haplotypes (600-1000)", lambda S: S.s4(16), (3, 12)),
    ("S5: 1024 windows of the ragged active-region stream", lambda S: next(S.s5_stream(1024, windows_per_batch=1024)), (1, 500, 1023)),
])
def test_full_size_other_configs(pkg, engine, oracle, name, make, spot):
    """BASELINE configs 2, 4 and 5 at full size: batch-split invariance (a region computed alone gives the
    same bits as inside the big batch, whatever jobs / chunks / lane partners the planner gave it) and an
    oracle check of whole regions."""
    b = make(pkg.synth)
    got = engine.compute(b)
    assert got.stats["n_cells"] == b.n_cells and got.stats["n_pairs"] == b.n_pairs
    if name.startswith("S4"):
        assert got.rescued.all()
    for g in spot:
        sub = b.slice_regions(g, g + 1)
        o = int(b.region_out_beg[g])
        mine = got.log10[o:o + sub.n_pairs]
        alone = engine.compute(sub)
        assert np.array_equal(alone.log10.view(np.uint64), mine.view(np.uint64)), f"{name}: region {g} alone"
        want = oracle.batch(sub, threads=16)
        resc = want["rescued"].astype(bool)
        assert np.array_equal(got.rescued[o:o + sub.n_pairs].astype(bool), resc)
        assert _maxerr(mine[~resc], want["log10"][~resc]) <= TOL32
        assert _maxerr(mine[resc], want["log10"][resc]) <= TOL64


def test_call_surface_compute_likelihoods(pkg, engine, golden):
    """hc::IntelPairHMM::compute_likelihoods semantics (cap at best-4.5, poorly modelled reads erased)
    against the reference's own outputs (Appendix A.2 + six regions)."""
    a2 = golden["kat_appendix_a"]["a2"]
    regions = [dict(a2, lik_bits=np.array(a2["lik"]).reshape(-1).view(np.uint64).tolist())] + golden["ref_region_filter"]["regions"]
    for reg in regions:
        lik, kept = engine.compute_likelihoods([h.encode() for h in reg["haps"]], [r.encode() for r in reg["reads"]],
                                               [q.encode() for q in reg["quals"]])
        assert kept.tolist() == [i for i, k in enumerate(reg["keep"]) if k]
        want = np.array(reg["lik_bits"], np.uint64).view(np.float64).reshape(lik.shape)
        assert _maxerr(lik, want) <= TOL32


def test_submit_wait_pipeline_and_staged_path(pkg, engine, oracle):
    batches = [pkg.synth.random_small(50 + i, n_regions=5, general_gaps=bool(i % 2)) for i in range(6)]
    sync = [engine.compute(b).log10 for b in batches]
    # two tickets in flight (pipeline depth 2), waited in order
    t0 = engine.submit(batches[0])
    for i in range(1, len(batches)):
        t1 = engine.submit(batches[i])
        assert np.array_equal(engine.wait(t0).log10.view(np.uint64), sync[i - 1].view(np.uint64))
        t0 = t1
    assert np.array_equal(engine.wait(t0).log10.view(np.uint64), sync[-1].view(np.uint64))
    # device-resident form gives the same bits as the host-buffer form
    st = engine.stage(batches[2])
    ms, launches = engine.run_staged(st, 2)
    res = engine.fetch_staged(st, batches[2].n_pairs)
    engine.free_staged(st)
    assert ms > 0 and launches > 0
    assert np.array_equal(res.log10.view(np.uint64), sync[2].view(np.uint64))


def test_run_staged_pipelined_gives_the_same_bits_as_compute(pkg, engine):
    """The path bench.py times (`value`): steps issued back to back, step i on staged batch i % n, every batch
    on its own stream so FP32 and FP64 launches of neighbouring steps overlap.  Every batch -- with and
    without FP64 redos, single- and multi-shape -- must come back with the bits phmm_compute gives."""
    batches = [pkg.synth.s3(3, seed=11), pkg.synth.s4(2, n_reads=24, n_haps=4, seed=12), pkg.synth.s5_batch(48, seed=13),
               pkg.synth.s2(6, seed=14), pkg.synth.random_small(15, n_regions=9, max_reads=30, max_haps=6)]
    want = [engine.compute(b) for b in batches]
    staged = [engine.stage(b) for b in batches]
    try:
        for steps in (len(batches), 3 * len(batches) + 2):           # the second round re-runs batches on their own streams
            ms, launches = engine.run_staged_pipelined(staged, steps)
            assert ms > 0 and launches >= 2 * steps
            for b, st, w in zip(batches, staged, want):
                got = engine.fetch_staged(st, b.n_pairs)
                assert np.array_equal(got.log10.view(np.uint64), w.log10.view(np.uint64))
                assert np.array_equal(got.rescued, w.rescued)
                assert got.stats["n_rescued"] == w.stats["n_rescued"]
    finally:
        for st in staged:
            engine.free_staged(st)


@pytest.mark.parametrize("name,make", [
    ("S4 long, all rescued", lambda s: s.s4(3, n_reads=24, n_haps=5)),
    ("S3: hardly any pair rescued", lambda s: s.s3(1)),
    ("mixed: good and hopeless reads, ragged", lambda s: s.random_small(91, n_regions=12, max_reads=30, max_haps=9, max_read_len=200,
                                                                         max_hap_len=300, general_gaps=False)),
    ("mixed, per-base gap penalties", lambda s: s.random_small(92, n_regions=8, max_reads=20, max_haps=6, max_read_len=200, max_hap_len=300)),
    ("every length", lambda s: s.random_small(93, n_regions=30, max_reads=12, max_haps=4, max_read_len=255, max_hap_len=97, general_gaps=False)),
])
def test_fp64_first_order_gives_identical_results(pkg, engine, oracle, name, make):
    """phmm_options.fp64_first = 2: every pair is scored in FP64 first and the FP32 pass only runs where the FP64 sum
    does not prove the underflow (phmm_kernels.cuh: kCertainUnderflow64).  log10 values and rescue decisions must be
    BIT-IDENTICAL to the FP32-first order, and within tolerance of the oracle."""
    b = make(pkg.synth)
    want = engine.compute(b)
    with pkg.PairHMMEngine(devices=[0], fp64_first=2) as eng:
        for _ in range(2):                                   # twice: the second batch also sees last_rescue_frac
            got = eng.compute(b)
            assert np.array_equal(got.rescued, want.rescued), name
            assert np.array_equal(got.log10.view(np.uint64), want.log10.view(np.uint64)), name
            assert got.stats["n_rescued"] == want.stats["n_rescued"]
            keep = ~want.rescued.astype(bool)                # raw FP32 of a scored pair is the same number
            assert np.array_equal(got.raw32[keep].view(np.uint32), want.raw32[keep].view(np.uint32)), name
            assert np.array_equal(got.raw64.view(np.uint64), want.raw64.view(np.uint64)), name
    check(want, oracle.batch(b, threads=16), what=name)


def test_fp64_first_kicks_in_after_a_dense_batch(pkg, oracle):
    """Auto mode: the order follows the device's previous batch (> 75% redone -> FP64 first), and falls back."""
    dense, sparse = pkg.synth.s4(2, n_reads=24, n_haps=4), pkg.synth.s3(1)
    with pkg.PairHMMEngine(devices=[0]) as eng:
        seq = [dense, dense, sparse, sparse, dense]
        launches = []
        for b in seq:
            got = eng.compute(b)
            check(got, oracle.batch(b, threads=16), what="auto order")
            launches.append(got.stats["kernel_launches"])
        assert launches[1] > launches[0]                     # second dense batch: FP64-first chain (one more launch pair)
        assert launches[3] < launches[2]                     # second sparse batch: back to FP32 first


def test_use_double_switch(pkg, engine, oracle):
    """phmm_options.use_double = the reference's g_use_double (intel_pairhmm.hpp:58,71,135): the FP32 pass is
    skipped (raw FP32 = 0.0f), every pair is scored in FP64 and reported as rescued."""
    b = pkg.Batch.concat([pkg.synth.random_small(21, n_regions=5, general_gaps=False), pkg.synth.s3(1, seed=3)])
    want = oracle.batch(b, threads=16)
    logd = np.log10(want_raw64(oracle, b)) - oracle.log10_init()[1]
    with pkg.PairHMMEngine(devices=[0], use_double=True) as eng:
        got = eng.compute(b)
    assert got.rescued.all() and got.stats["n_rescued"] == b.n_pairs
    assert not got.raw32.any()
    assert _maxerr(got.log10, logd) <= TOL64
    # where the FP32 path would NOT have been rescued the two precisions agree to FP32 accuracy
    keep = ~want["rescued"].astype(bool)
    assert _maxerr(got.log10[keep], want["log10"][keep]) <= 1e-4


def want_raw64(oracle, b):
    """FP64 forward sum of EVERY pair from the oracle (its batch() only keeps them for rescued pairs)."""
    out = np.empty(b.n_pairs, np.float64)
    ob = b.region_out_beg
    for g in range(b.n_regions):
        r0, r1 = int(b.region_read_beg[g]), int(b.region_read_beg[g + 1])
        h0, h1 = int(b.region_hap_beg[g]), int(b.region_hap_beg[g + 1])
        k = int(ob[g])
        for r in range(r0, r1):
            a, z = int(b.read_off[r]), int(b.read_off[r + 1])
            for h in range(h0, h1):
                c, d = int(b.hap_off[h]), int(b.hap_off[h + 1])
                out[k] = oracle.forward(b.read_bases[a:z], b.read_q[a:z], b.read_i[a:z], b.read_d[a:z], b.read_c[a:z],
                                        b.hap_bases[c:d], "f64")
                k += 1
    return out


def test_pinned_inputs_upload_without_the_staging_copy(pkg, engine):
    """PHMM_BATCH_PINNED_INPUTS: the byte arrays go to the device straight from the caller's page-locked memory;
    same bits as the staged-copy path, through compute and through several tickets in flight, with constant
    and with per-base gap penalties."""
    for b in (pkg.synth.s3(3), pkg.synth.random_small(5, n_regions=9, max_reads=30, max_haps=6),
              next(pkg.synth.s5_stream(64, windows_per_batch=64))):
        want = engine.compute(b)
        b.pin()
        try:
            got = engine.compute(b)
            assert np.array_equal(got.log10.view(np.uint64), want.log10.view(np.uint64))
            assert np.array_equal(got.rescued, want.rescued)
            tickets = [engine.submit(b) for _ in range(2)]
            for t in tickets:
                assert np.array_equal(engine.wait(t).log10.view(np.uint64), want.log10.view(np.uint64))
        finally:
            b.unpin()
        assert np.array_equal(engine.compute(b).log10.view(np.uint64), want.log10.view(np.uint64))


def test_edge_cases_and_errors(pkg, engine, oracle):
    B = pkg.Batch
    # regions without reads or without haplotypes contribute no pairs; empty batch is fine
    b = B.from_regions([([], [], [b"ACGT"]), ([b"ACGT"], [b"FFFF"], []), ([b"ACGT"], [b"FFFF"], [b"ACGT"])])
    got = engine.compute(b)
    assert got.log10.size == 1 and abs(got.log10[0] - oracle.batch(b)["log10"][0]) <= TOL32
    assert engine.compute(B.from_regions([])).log10.size == 0
    # one-base read, one-base haplotype
    b = B.from_regions([([b"A", b"C"], [b"I", b"#"], [b"A", b"N", b"ACGTACGT"])])
    check(engine.compute(b), oracle.batch(b), what="1-base")
    # longest supported read (255) against a long haplotype
    rng = np.random.default_rng(9)
    hap = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 3000)]
    rd = hap[100:355].copy()
    b = B.from_regions([([rd], [np.full(255, 63, np.uint8)], [hap, hap[:700]])])
    check(engine.compute(b), oracle.batch(b), what="R=255 H=3000")
    # errors: read too long -> UNSUPPORTED (5); empty read -> INVALID_ARG (1); bad ticket (6)
    with pytest.raises(pkg.PhmmError) as ei:
        engine.compute(B.from_regions([([np.full(2049, 65, np.uint8)], [np.full(2049, 70, np.uint8)], [b"ACGT"])]))
    assert ei.value.code == 5
    with pytest.raises(pkg.PhmmError) as ei:
        engine.compute(B.from_regions([([b""], [b""], [b"ACGT"])]))
    assert ei.value.code == 1
    with pytest.raises(pkg.PhmmError) as ei:
        engine.wait(123456)
    assert ei.value.code == 6
    # the engine is still healthy afterwards
    b = B.from_regions([([b"ACGT"], [b"FFFF"], [b"ACGT"])])
    check(engine.compute(b), oracle.batch(b), what="after errors")


def test_hunt_for_a_fast_vs_exact_rescue_flip(pkg, engine, exact_engine):
    """The default engine contracts mul+add into FMA; its raw FP32 sum can differ from the reference's (= the exact
    engine's, bit for bit) in the last place, so a pair whose raw sum lands within an ulp or two of MIN_ACCEPTED
    (1e-28f, pairhmm_common.h:16) can take the other side of `raw < 1e-28f` (intel_pairhmm.hpp:137).  Hunt for such
    pairs in two stages: (1) 1.5 million short reads whose likelihoods straddle the threshold by orders of magnitude;
    (2) the read that came closest, with the qualities of its MATCHING bases redrawn 1.5 million times -- each of those
    moves the likelihood by up to 1e-4 relative in steps far below a float ulp, which paves the neighbourhood of the
    threshold at hundreds of pairs per ulp.  Whatever the number of flips, a flip may move the final log10 only by the FP32-vs-FP64
    difference (<= 1e-4): the parity bar of north_star holds; BIT identity of the VCF is what exact_fp32 is for
    (INTEGRATION.md)."""
    rng = np.random.default_rng(2024)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    n, R, H = 1_500_000, 24, 40
    hap = alpha[rng.integers(0, 4, H)]
    thr = np.float32(1e-28)

    def run(reads, quals):
        b = pkg.Batch([0, n], [0, 1], np.arange(n + 1) * R, reads.reshape(-1), quals.reshape(-1), [0, H], hap)
        fast, exact = engine.compute(b), exact_engine.compute(b)
        ulps = np.abs(exact.raw32.view(np.int32).astype(np.int64) - thr.view(np.int32).astype(np.int64))
        flips = fast.rescued != exact.rescued
        d = np.abs(fast.log10 - exact.log10)
        assert np.nanmax(d) <= TOL32                                   # a flip never costs more than FP32-vs-FP64
        assert (ulps[flips] <= 16).all()                               # and flips only ever happen AT the threshold (2e-6 relative)
        assert np.array_equal(fast.rescued[~flips], exact.rescued[~flips])
        return fast, exact, ulps, flips

    # stage 1: 8..12 mismatches on every other base (contiguous ones would be absorbed by one cheap insertion), random
    # qualities: log10 L spreads over about [-70, -47] around the threshold at -64.1 = log10(1e-28 / 2^120)
    reads = np.tile(hap[8:8 + R], (n, 1))
    n_mm = rng.integers(8, 13, n)
    quals = (33 + rng.integers(20, 41, (n, R))).astype(np.uint8)
    for k, i in enumerate(range(1, R, 2)):
        hit = (k < n_mm)
        reads[hit, i] = alpha[(np.searchsorted(alpha, reads[hit, i]) + 1) % 4]
    fast, exact, ulps, flips = run(reads, quals)
    frac_below = float((exact.raw32 < thr).mean())
    assert 0.05 < frac_below < 0.95                                    # the hunt really straddles the threshold
    best = int(np.argmin(ulps))
    print(f"\nrescue-flip hunt, stage 1: {n} pairs, {frac_below:.1%} below 1e-28f, closest {int(ulps[best])} ulp away, "
          f"{int(flips.sum())} flips; raw FP32 differs in {int((fast.raw32 != exact.raw32).sum())} pairs")
    # stage 2: that read, about half of its matching-base qualities redrawn, the closest variant taken as the next
    # centre, the quality band narrowed each round (a match prior is 1 - 10^(-q/10): ever finer steps)
    n = 400_000
    rd, q = reads[best].copy(), quals[best].copy()
    matching = np.array([i for i in range(R) if not (i % 2 == 1 and (i // 2) < n_mm[best])])
    total_near = total_flips = 0
    for band_lo in (20, 27, 32, 36):
        reads2, quals2 = np.tile(rd, (n, 1)), np.tile(q, (n, 1))
        redraw = rng.random((n, len(matching))) < 0.5
        newq = (33 + rng.integers(band_lo, 41, (n, len(matching)))).astype(np.uint8)
        quals2[:, matching] = np.where(redraw, newq, quals2[:, matching])
        quals2[0] = q
        fast2, exact2, ulps2, flips2 = run(reads2, quals2)
        q = quals2[int(np.argmin(ulps2))].copy()
        near = ulps2 <= 2
        print(f"rescue-flip hunt, stage 2, quality band [{band_lo}, 40]: {n} variants, {int(near.sum())} within 2 ulp of 1e-28f "
              f"({int((ulps2 <= 8).sum())} within 8), {int(flips2.sum())} fast-vs-exact rescue flips, "
              f"{int((fast2.raw32 != exact2.raw32).sum())} raw FP32 sums differ in the last place(s)")
        total_near += int(near.sum()); total_flips += int(flips2.sum())
    print(f"rescue-flip hunt: {total_near} pairs within 2 ulp of the threshold examined, {total_flips} rescue decisions flipped by FMA contraction")
    assert total_near >= 20                                            # the hunt really reached the threshold


@pytest.mark.parametrize("name,make", [
    ("S3", lambda s: s.s3(2)),
    ("S2 (packed lane groups)", lambda s: s.s2(8)),
    ("S5 ragged windows", lambda s: s.s5_batch(48, seed=5)),
    ("S4 all rescued", lambda s: s.s4(2, n_reads=24, n_haps=4)),
    ("tiny haplotypes, every length", lambda s: s.random_small(71, n_regions=30, max_reads=12, max_haps=4, max_read_len=255, max_hap_len=40, general_gaps=False)),
])
def test_scaled_and_reference_order_recurrences_agree(pkg, engine, oracle, name, make):
    """The default engine runs the SCALED recurrence for constant gap penalties with i == d (phmm_kernels.cuh MODE 3:
    six FP32-pipe instructions per cell); phmm_options.recurrence = 1 keeps the reference's operation order
    (FMA-contracted, MODE 2).  Both must meet the oracle's bar, take the same rescue decisions on these inputs, and
    agree with each other far inside it; haplotypes of a few bases exercise the per-haplotype scale split (row 0 of the
    scaled recurrence must neither overflow nor push M towards the underflow guard)."""
    b = make(pkg.synth)
    want = oracle.batch(b, threads=16)
    got = engine.compute(b)
    check(got, want, what=name + " scaled")
    with pkg.PairHMMEngine(devices=[0], recurrence=1) as ref_order:
        other = ref_order.compute(b)
    check(other, want, what=name + " reference order")
    keep = ~want["rescued"].astype(bool)
    assert _maxerr(got.log10[keep], other.log10[keep]) <= 2e-5
    assert _maxerr(got.log10[~keep], other.log10[~keep]) <= 1e-9


@pytest.mark.parametrize("i_range,c_range", [((10, 21), (10, 30)),     # MODE 4 (folded): spread 10 <= c_min, i_min = 10
                                             ((4, 15), (10, 30)),      # i_min < 10 (pMM down to 0.2): MODE 0
                                             ((0, 4), (10, 30)),       # pMM = 0 rows: MODE 0
                                             ((30, 70), (10, 30)),     # spread 39 > c_min + 10: MODE 0
                                             ((40, 60), (25, 60))])    # MODE 4, wide
def test_per_base_gap_penalties_either_side_of_the_scaled_mode_guards(pkg, engine, oracle, i_range, c_range):
    """MODE 4 (scaled + folded, per-base gap penalties) is selected only where the engine's guards hold
    (phmm_engine.cu: scaled_general_is_safe -- continuation >= Q10, gap-open spread <= c_min + 10, gap-open >= Q10
    because the row's weights are divided by its pMM); everything else runs MODE 0.  Parity on both sides."""
    b = pkg.synth.random_small(77, n_regions=16, max_reads=40, max_haps=8, max_read_len=200, max_hap_len=300, general_gaps=True)
    assert b.explicit_gaps
    rng = np.random.default_rng(i_range[0] * 131 + c_range[0])
    n = len(b.read_i)
    b.read_i[:] = rng.integers(i_range[0], i_range[1], n)
    b.read_d[:] = rng.integers(i_range[0], i_range[1], n)
    b.read_c[:] = rng.integers(c_range[0], c_range[1], n)
    check(engine.compute(b), oracle.batch(b, threads=8), what=f"per-base gaps i,d in {i_range}, c in {c_range}")


@pytest.mark.parametrize("gap", [(40, 40, 35), (96, 96, 43), (100, 100, 43), (127, 127, 60), (33, 33, 43), (73, 73, 12),
                                 (10, 10, 30), (9, 9, 30), (3, 3, 20)])     # Q10: the last folded one; below: reference order (pMM -> 0)
def test_scaled_recurrence_over_the_range_of_constant_gap_penalties(pkg, engine, oracle, gap):
    """The scale split of the scaled recurrence depends on the gap-open probability (row 0 must not overflow, M^ must
    not underflow): from a cheap gap (Q40 raw byte: s = 1) over the reference's own 'I' to penalties beyond Q96, where
    the engine keeps the reference-order kernels; short and long haplotypes."""
    rng = np.random.default_rng(sum(gap))
    alpha = np.frombuffer(b"ACGT", np.uint8)
    regions = []
    for H in (3, 40, 700):
        hap = alpha[rng.integers(0, 4, H)]
        reads, quals = [], []
        for rl in (1, 7, 33, 100, 150, 151, 255):
            o = int(rng.integers(0, max(1, H - rl + 1)))
            r = np.resize(hap[o:o + rl], rl).copy()
            m = rng.random(rl) < 0.05; r[m] = alpha[rng.integers(0, 4, int(m.sum()))]
            reads.append(r); quals.append((33 + rng.integers(2, 42, rl)).astype(np.uint8))
        regions.append((reads, quals, [hap, hap[::-1].copy()]))
    b0 = pkg.Batch.from_regions(regions)
    b = pkg.Batch(b0.region_read_beg, b0.region_hap_beg, b0.read_off, b0.read_bases, b0.read_q, b0.hap_off, b0.hap_bases,
                  gap_open_i=gap[0], gap_open_d=gap[1], gap_cont_c=gap[2])
    check(engine.compute(b), oracle.batch(b, threads=8), what=f"gap bytes {gap}")
