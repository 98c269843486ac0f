"""Smith-Waterman haplotype -> reference alignment: hc::IntelSWAligner::align
(smithwaterman/intel_smithwaterman.hpp:29-44, native/PairWiseSW.h) rebuilt as a batched sm_100a kernel
behind phmm_sw_align.  Integer work: offsets and CIGARs must be IDENTICAL.

  CPU : the plain-C restatement (oracle/sw_oracle.c) against the committed golden fixture made by the
        compiled reference, and against the compiled reference itself when oracle/_ref is present;
  GPU : phmm_sw_align against the fixture and against the oracle on seeded pairs, error behaviour, the C++
        mirror hc::B200SWAligner inside the reference's whole driver (byte-identical VCF).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from _sw_cases import sw_cases, PARAMS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ARGS = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]


def _oracle_align(lib, fn, ref, alt, prm):
    buf = C.create_string_buffer(16384)
    off = getattr(lib, fn)(ref, len(ref), alt, len(alt), *prm, buf, 16384)
    return off, buf.value.decode()


@pytest.fixture(scope="module")
def sw_oracle(oracle):
    lib = oracle.lib
    lib.sw_oracle_align.argtypes = _ARGS
    return lambda ref, alt, prm: _oracle_align(lib, "sw_oracle_align", ref, alt, prm)


@pytest.fixture(scope="module")
def sw_golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_sw_pairs.json")))


def test_sw_oracle_matches_the_reference_fixture(sw_oracle, sw_golden):
    cases = sw_cases(sw_golden["seed"], sw_golden["n"])
    assert sw_golden["params"] == PARAMS
    for k, ((ref, alt), want) in enumerate(zip(cases, sw_golden["results"])):
        assert list(sw_oracle(ref, alt, PARAMS[k % 4])) == want, f"case {k}"
    kinds = "".join(c for _, cig in sw_golden["results"] for c in cig if c.isalpha())
    assert all(op in kinds for op in "MIDS")                              # the fixture exercises every operator


def test_sw_oracle_matches_the_compiled_reference(sw_oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ref.lib.ref_sw_align.argtypes = _ARGS
    for k, (r, a) in enumerate(sw_cases(7, 400, max_len=400)):
        assert sw_oracle(r, a, PARAMS[k % 4]) == _oracle_align(ref.lib, "ref_sw_align", r, a, PARAMS[k % 4]), f"case {k}"


@pytest.mark.gpu
def test_sw_kernel_matches_fixture_and_oracle(pkg, sw_oracle, sw_golden):
    cases = sw_cases(sw_golden["seed"], sw_golden["n"])
    for p in range(4):                                                   # one launch per parameter set
        idx = [k for k in range(len(cases)) if k % 4 == p]
        got, ms = pkg.sw_align([cases[k] for k in idx], params=tuple(PARAMS[p]))
        assert [list(g) for g in got] == [sw_golden["results"][k] for k in idx]
    big = sw_cases(99, 1500, max_len=700)                                # every columns-per-lane class, one launch
    got, ms = pkg.sw_align(big)
    assert ms > 0
    for k, ((r, a), g) in enumerate(zip(big, got)):
        assert g == sw_oracle(r, a, PARAMS[0]), f"case {k}: ref {len(r)} alt {len(a)}"


@pytest.mark.gpu
def test_sw_edge_cases_and_errors(pkg, sw_oracle):
    ref = b"ACGTACGTAC"
    for alt in (b"ACGTACGTAC", b"ACGTTCGTAC", b"A", b"TTTTTTTTTTTTTTTT", ref * 60, ref[3:8]):
        assert pkg.sw_align([(ref, alt)])[0][0] == sw_oracle(ref, alt, PARAMS[0])
        assert pkg.sw_align([(alt, ref)])[0][0] == sw_oracle(alt, ref, PARAMS[0])
    assert pkg.sw_align([])[0] == []
    with pytest.raises(pkg.PhmmError) as ei:
        pkg.sw_align([(b"A" * 1024, b"ACGT")])                           # beyond what the reference's arrays hold
    assert ei.value.code == 5
    with pytest.raises(pkg.PhmmError) as ei:
        pkg.sw_align([(b"", b"ACGT")])                                   # the reference throws invalid_argument
    assert ei.value.code == 1
    rng = np.random.default_rng(3)
    r = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 800)]
    a = np.delete(r, np.arange(5, 800, 9)).tobytes()                       # a deletion every 9 bases: ~180 CIGAR elements
    r = r.tobytes()
    off, cig = pkg.sw_align([(r, a)])[0][0]
    assert (off, cig) == sw_oracle(r, a, PARAMS[0]) and cig.count("D") > 50
    assert pkg.sw_align([(r, a)] * 3)[0] == [(off, cig)] * 3                # default room too small: the wrapper retries
    with pytest.raises(pkg.PhmmError) as ei:
        pkg.sw_align([(r, a)], cap_elems=8)
    assert ei.value.code == 5


@pytest.mark.gpu
def test_sw_aligner_inside_the_reference_driver(pkg, tmp_path):
    """hc::B200SWAligner (include/b200_smithwaterman.hpp) swapped in for hc::IntelSWAligner inside the
    reference's assembler, together with the B200 PairHMM engine: the VCF stays byte-identical."""
    exe = os.path.join(ROOT, "oracle", "_ref", "hc_e2e_b200_sw")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/hc_e2e_b200_sw not built")
    prefix = str(tmp_path / "chrm_like")
    pkg.synth.chrm_like(prefix)
    out = str(tmp_path / "sw.vcf")
    r = subprocess.run([exe, "-I", prefix + ".sam", "-R", prefix + ".fa", "-O", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(out).read() == open(os.path.join(ROOT, "tests", "golden", "chrm_like.ref.vcf")).read()
