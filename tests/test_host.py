"""CPU suite, part 2: host logic and the C-ABI library surface (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "phmm.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+char\*|int64_t|int|void)\s+(phmm_\w+)\s*\(", hdr, re.M))
    assert declared >= {"phmm_create", "phmm_destroy", "phmm_compute", "phmm_submit", "phmm_wait", "phmm_strerror"}
    L = pkg.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"libphmm_b200.so does not export {name}"
    assert declared == set(pkg.EXPORTS)
    assert L.phmm_abi_version() == 1


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/phmm.h must compile as C99 (what cgo / JNI / ctypes-style bindings see), and a C
    translation unit that names every entry point must link against the library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "phmm.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    names = re.findall(r"^\s*(?:const\s+char\*|int64_t|int|void)\s+(phmm_\w+)\s*\(", open(hdr).read(), re.M)
    src = tmp_path / "link_all.c"
    src.write_text('#include "phmm.h"\n#include <stdio.h>\nint main(void) {\n  void* f[] = {' +
                   ", ".join(f"(void*)(size_t){n}" for n in names) +
                   '};\n  printf("%d %d\\n", (int)(sizeof f / sizeof f[0]), phmm_abi_version());\n  return 0;\n}\n')
    libdir = os.path.join(ROOT, "gatk-haplotypecaller-cpp17_b200")
    exe = tmp_path / "link_all"
    subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L" + libdir, "-lphmm_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == len(names) and int(out[1]) == 1


def test_only_sm100a_code_in_the_library(pkg):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_strerror_and_no_cpu_fallback(pkg):
    L = pkg.lib()
    assert L.phmm_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.phmm_strerror(2)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.PhmmError) as ei:       # creating an engine must fail loudly, never fall back
        pkg.PairHMMEngine()
    assert ei.value.code == 2


def test_batch_container(pkg):
    b = pkg.Batch.from_regions([([b"ACGT", b"AC"], [b"FFFF", b"FF"], [b"ACGTA", b"AC", b"A"]),
                                ([b"TTT"], [b"III"], [b"TTTT"])])
    assert (b.n_regions, b.n_reads, b.n_haps, b.n_pairs) == (2, 3, 4, 7)
    assert b.n_cells == (4 + 2) * (5 + 2 + 1) + 3 * 4
    assert b.region_out_beg.tolist() == [0, 6, 7]
    assert not b.explicit_gaps and b.read_i.tolist() == [ord("I")] * 9 and b.read_c.tolist() == [ord("+")] * 9
    s = b.slice_regions(1, 2)
    assert (s.n_regions, s.n_reads, s.n_haps) == (1, 1, 1) and bytes(s.read_bases) == b"TTT" and bytes(s.hap_bases) == b"TTTT"
    cs = b.c_struct()
    assert cs.n_reads == 3 and not cs.read_i and cs.gap_open_i == ord("I") and cs.gap_cont_c == ord("+")


def test_synthetic_generators_match_the_named_shapes(pkg):
    s2, s3 = pkg.synth.s2(3), pkg.synth.s3(2)
    assert s2.reads_per_region.tolist() == [64] * 3 and s2.haps_per_region.tolist() == [8] * 3
    assert set(np.diff(s2.read_off)) == {100} and set(np.diff(s2.hap_off)) == {300}
    assert s3.reads_per_region.tolist() == [256] * 2 and s3.haps_per_region.tolist() == [16] * 2
    assert set(np.diff(s3.read_off)) == {150} and set(np.diff(s3.hap_off)) == {500}
    assert s3.n_cells == 2 * 256 * 16 * 150 * 500
    assert pkg.synth.s3(2).read_bases.tobytes() == s3.read_bases.tobytes()          # seeded
    q = s3.read_q.astype(int) - 33
    assert q.min() >= 20 and q.max() <= 40
    s4 = pkg.synth.s4(2, n_reads=5, n_haps=3)
    rl, hl = np.diff(s4.read_off), np.diff(s4.hap_off)
    assert rl.min() >= 150 and rl.max() <= 250 and hl.min() >= 600 and hl.max() <= 1000
    g = pkg.synth.s3(1, general_gaps=True)
    assert g.explicit_gaps and len(set(g.read_i.tolist())) > 1
    tot = sum(b.n_regions for b in pkg.synth.s5_stream(20, windows_per_batch=8))
    assert tot == 20


def test_shard_bounds_balance_and_cover(pkg):
    b = pkg.synth.random_small(5, n_regions=23)
    cells = b.region_cells()
    assert int(cells.sum()) == b.n_cells
    for world in (1, 2, 3, 4, 8):
        cut = pkg.shard_bounds(cells, world)
        assert cut[0] == 0 and cut[-1] == b.n_regions and all(x <= y for x, y in zip(cut, cut[1:]))
        offs = [pkg.shard_regions(b, r, world)[1] for r in range(world)]
        sizes = [pkg.shard_regions(b, r, world)[0].n_pairs for r in range(world)]
        assert offs == np.concatenate([[0], np.cumsum(sizes)])[:-1].tolist() and sum(sizes) == b.n_pairs
    big = pkg.synth.s2(64)
    per = [pkg.shard_regions(big, r, 8)[0].n_cells for r in range(8)]
    assert max(per) == min(per)                     # equal regions -> exact balance


def test_normalize_filter_semantics(pkg):
    # intel_pairhmm.hpp:24-46: cap at best-4.5; drop when best < min(2, ceil(0.02 len)) * -4
    lik = np.array([[-1.0, -9.0, -5.5], [-9.0, -20.0, -8.5], [-4.0, -4.0, -4.0]])
    keep = pkg.normalize_filter(lik, np.array([100, 100, 10], np.int32))
    assert lik[0].tolist() == [-1.0, -5.5, -5.5]
    assert keep.tolist() == [1, 0, 1]               # thresholds: -8, -8, -4 (best == threshold is kept)
    assert lik[1].tolist() == [-9.0, -13.0, -8.5]


def _check_plan(pkg, b, host_threads=1):
    """Invariants of the planner (phmm_plan, pure host logic): every read of at most 255 bases sits in
    exactly one warp job of its own region, whose lane-group shape holds the job's longest read with a dummy
    row to spare; lane-aligned jobs only hold lane-aligned reads; longer reads go to the long-read kernel,
    one pair per haplotype of their region; cells and pairs are counted once."""
    info, jobs = pkg.plan(b, host_threads=host_threads)
    rl = np.diff(b.read_off)
    region_of_read = np.repeat(np.arange(len(b.region_read_beg) - 1), np.diff(b.region_read_beg))
    nh = np.diff(b.region_hap_beg)
    assert info["n_pairs"] == b.n_pairs and info["n_cells"] == b.n_cells
    assert info["n_jobs"] == len(jobs) == sum(info["jobs_ragged"]) + sum(info["jobs_aligned"])
    seen = np.zeros(len(rl), np.int32)
    n = info["n_shapes"]
    for row in jobs:
        slot, region, reads = int(row[0]), int(row[1]), [int(r) for r in row[2:] if r >= 0]
        G, K = info["shapes"][slot % n]
        assert all(region_of_read[r] == region for r in reads)
        if slot % n >= n - 3:                                              # PACKED (the last three shapes): free-width groups of nl lanes
            assert slot >= n and len({int(rl[r]) for r in reads}) == 1 and rl[reads[0]] % K == 0
            nl = int(rl[reads[0]]) // K
            assert nl in (8, 10, 15, 16) and 1 <= len(reads) <= 2 * min(32 // nl, 4)
            seen[reads] += 1
            continue
        assert 1 <= len(reads) <= 2 * (32 // G)
        assert max(rl[r] for r in reads) + 1 <= K * G                      # one dummy row on top is mandatory
        if slot >= n:                                                      # lane-aligned job
            assert info["mode"] != 0
            assert all(rl[r] % K == 0 and K * G - rl[r] >= K for r in reads)
        seen[reads] += 1
    short = rl <= 255
    has_haps = nh[region_of_read] > 0
    assert np.array_equal(seen[short & has_haps], np.ones(int((short & has_haps).sum()), np.int32))
    assert seen[~short].sum() == 0 and seen[~has_haps].sum() == 0
    assert info["n_long_pairs"] == int(nh[region_of_read[~short]].sum())
    assert 1 <= info["haps_per_job"] and info["haps_per_job"] * info["hap_chunks"] >= nh.max()
    assert info["haps_per_job64"] * info["hap_chunks64"] >= nh.max()
    return info, jobs


def test_planner_invariants(pkg):
    S = pkg.synth
    _check_plan(pkg, S.random_small(11, n_regions=20, max_reads=40, max_haps=9, max_read_len=255, general_gaps=False))
    info, _ = _check_plan(pkg, S.random_small(12, n_regions=6))                       # per-base gap penalties
    assert info["mode"] == 0 and sum(info["jobs_aligned"]) == 0
    info, jobs = _check_plan(pkg, S.s3(2))
    assert info["mode"] == 2 and info["n_jobs"] == 2 * 64 and sum(info["jobs_aligned"]) == 128   # 150 = 15 lanes x 10 rows
    info, jobs = _check_plan(pkg, S.s2(16))                                # 100 = 10 lanes x 10 rows: three groups per warp
    assert info["jobs_aligned"][15] == 16 * 11 and info["n_jobs"] == 16 * 11 and info["shapes"][15] == (32, 10)
    b = next(S.s5_stream(128, windows_per_batch=128))
    one, jobs1 = _check_plan(pkg, b, host_threads=1)
    four, jobs4 = _check_plan(pkg, b, host_threads=4)
    assert np.array_equal(jobs1, jobs4) and one == four                    # the plan does not depend on the thread count
    # sorted packing: at most one partial job per region and shape class boundary -> close to reads / 4
    n_reads = len(b.read_off) - 1
    assert one["n_jobs"] <= n_reads / 4 + 2 * 128
    # reads beyond 255 bases are planned for the long-read kernel
    rng = np.random.default_rng(5)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    mk = lambda n: alpha[rng.integers(0, 4, n)]
    reads = [mk(300), mk(100), mk(2048), mk(255), mk(256)]
    b = pkg.Batch.from_regions([(reads, [np.full(len(r), 60, np.uint8) for r in reads], [mk(500), mk(400), mk(30)]),
                                ([mk(50)], [np.full(50, 60, np.uint8)], [])])
    info, jobs = _check_plan(pkg, b)
    assert info["n_long_pairs"] == 3 * 3 and info["n_jobs"] == 1


def test_planner_chunks_follow_the_makespan_model(pkg):
    """Many haplotypes per region and plenty of jobs -> several haplotypes per (job, chunk) unit; a ragged
    stream with a few many-haplotype regions -> short units, so that no warp runs far longer than the rest."""
    S = pkg.synth
    info, _ = pkg.plan(S.s3(128))
    assert info["haps_per_job"] == 4 and info["hap_chunks"] == 4 and info["haps_per_job64"] == 1
    info, _ = pkg.plan(next(S.s5_stream(512, windows_per_batch=512)))
    assert info["haps_per_job"] <= 3
    info, _ = pkg.plan(S.s3(1))
    assert info["haps_per_job"] == 1 and info["hap_chunks"] == 16          # few jobs: every haplotype its own unit


def test_argument_checks_without_a_device(pkg):
    """phmm_validate: what phmm_submit / phmm_submit_gl refuse before touching a device (include/phmm.h error codes)."""
    S, B = pkg.synth, pkg.Batch
    b = S.random_small(3, n_regions=4, max_reads=6, max_haps=3, general_gaps=False)
    assert pkg.validate(b) == 0
    assert pkg.validate(B.from_regions([([b""], [b""], [b"ACGT"])])) == 1                      # empty read
    assert pkg.validate(B.from_regions([([np.full(2049, 65, np.uint8)], [np.full(2049, 70, np.uint8)], [b"ACGT"])])) == 5
    nh, nr = b.haps_per_region, b.reads_per_region
    ok = [(g, 2, np.zeros(int(nh[g]), np.uint8), np.ones(int(nr[g]), np.uint8)) for g in range(b.n_regions)]
    assert pkg.validate(b, pkg.Sites(b, ok)) == 0
    assert pkg.validate(b, pkg.Sites(b, ok[::-1])) == 1                                          # regions must be non-decreasing
    bad_allele = [(0, 2, np.full(int(nh[0]), 2, np.uint8), None)]
    assert pkg.validate(b, pkg.Sites(b, bad_allele)) == 1                                        # allele 2 of a 2-allele site
    too_many = [(0, 8, np.zeros(int(nh[0]), np.uint8), None)]
    assert pkg.validate(b, pkg.Sites(b, too_many)) == 1                                          # > PHMM_MAX_ALLELES
    assert pkg.validate(b, pkg.Sites(b, [])) == 0
