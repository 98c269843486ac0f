"""The C++17 host mirror (include/b200_pairhmm.hpp): hc::B200PairHMM::compute_likelihoods with the
reference's call signature (pairhmm/intel_pairhmm.hpp:48-56), over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gatk-haplotypecaller-cpp17_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory, pkg):
    out = str(tmp_path_factory.mktemp("cpp") / "host_mirror_main")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "host_mirror_main.cpp"), "-o", out,
                    "-L" + PKG, "-lphmm_b200", "-Wl,-rpath," + PKG], check=True)
    return out


def _region_file(path, reg):
    with open(path, "w") as f:
        for h in reg["haps"]:
            f.write(f"H {h}\n")
        for r, q in zip(reg["reads"], reg["quals"]):
            f.write(f"R {r} {q}\n")


def test_header_compiles_and_fails_loudly_without_gpu(exe, tmp_path, golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = tmp_path / "r.txt"
    _region_file(p, golden["kat_appendix_a"]["a2"])
    r = subprocess.run([exe, str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr       # never a silent CPU path


@pytest.mark.gpu
def test_cpp_mirror_matches_reference_call_surface(exe, tmp_path, golden):
    a2 = golden["kat_appendix_a"]["a2"]
    regions = [dict(a2, lik_bits=np.array(a2["lik"]).reshape(-1).view(np.uint64).tolist())] + golden["ref_region_filter"]["regions"]
    for i, reg in enumerate(regions):
        p = tmp_path / f"r{i}.txt"
        _region_file(p, reg)
        r = subprocess.run([exe, str(p)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        kept = [int(x) for x in lines[0].split()[1:]]
        assert kept == [k for k, f in enumerate(reg["keep"]) if f]            # reads erased like :40-45
        got = np.array([[float(x) for x in ln.split()] for ln in lines[1:]]).reshape(len(kept), -1)
        want = np.array(reg["lik_bits"], np.uint64).view(np.float64).reshape(got.shape)
        assert np.abs(got - want).max() <= 1e-4


@pytest.mark.gpu
def test_region_batcher_equals_per_region_calls(exe, tmp_path, golden):
    """hc::B200RegionBatcher (cross-window batching): regions added one by one, flushed into
    several asynchronous batches, taken in reverse order -> the same kept reads and the same matrix
    (to the bit) as one hc::B200PairHMM::compute_likelihoods call per region."""
    a2 = golden["kat_appendix_a"]["a2"]
    regions = [a2] + golden["ref_region_filter"]["regions"]
    regions = regions + regions[::-1]
    paths = []
    for i, reg in enumerate(regions):
        p = tmp_path / f"b{i}.txt"
        _region_file(p, reg)
        paths.append(str(p))
    single = []
    for p in paths:
        r = subprocess.run([exe, p], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        single.append([ln.rstrip() for ln in r.stdout.strip().splitlines()])
    r = subprocess.run([exe, "--batched"] + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert int(r.stderr.strip().split()[-1]) >= 2                       # really several batches
    blocks, cur = {}, None
    for ln in r.stdout.strip().splitlines():
        if ln.startswith("region "):
            cur = int(ln.split()[1]); blocks[cur] = []
        else:
            blocks[cur].append(ln.rstrip())
    assert sorted(blocks) == list(range(len(paths)))
    for i in range(len(paths)):
        assert blocks[i] == single[i], f"region {i}"


@pytest.mark.gpu
def test_cpp_sw_aligner_mirror(exe, pkg):
    """hc::B200SWAligner (include/b200_smithwaterman.hpp): align() and align_batch() give what the Python face
    of phmm_sw_align gives (which tests/test_sw.py pins to the reference)."""
    rng = np.random.default_rng(21)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    ref = alpha[rng.integers(0, 4, 300)].tobytes().decode()
    alts = [ref[20:200], ref[:100] + "ACGTAC" + ref[100:], ref[:150] + ref[161:], ref]
    want, _ = pkg.sw_align([(ref.encode(), a.encode()) for a in alts])
    r = subprocess.run([exe, "--sw", ref] + alts, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [(int(ln.split()[0]), ln.split()[1]) for ln in r.stdout.strip().splitlines()]
    assert got == want
    r = subprocess.run([exe, "--sw", ref, alts[1]], capture_output=True, text=True)
    assert r.returncode == 0 and (int(r.stdout.split()[0]), r.stdout.split()[1]) == want[1]


def test_cpp_sw_aligner_fails_loudly_without_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "--sw", "ACGTACGTAA", "ACGTTTACGTAA"], capture_output=True, text=True)
    assert r.returncode == 1 and "no usable sm_100" in r.stderr
