"""chrM-style END-TO-END parity (BASELINE config 1 / SURVEY.md section 8f-1): the reference's whole
driver (FASTA+SAM load, windowing, filters, clipping, local assembly, PairHMM, genotyping, VCF),
compiled unmodified from /root/reference with a Boost.Graph shim (oracle/hc_e2e.cpp), once around
hc::IntelPairHMM and once around hc::B200PairHMM.  Input: a synthetic chrM-like contig with at most one
read per start position (the reference's only source of non-determinism), regenerated from a seed.

  CPU : the reference-engine run reproduces the committed golden VCF;
  GPU : the B200-engine run writes a BIT-IDENTICAL VCF file; wall times of both are printed.
"""
import os
import subprocess
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "hc_e2e_ref")
B200_EXE = os.path.join(ROOT, "oracle", "_ref", "hc_e2e_b200")
BATCHED_EXE = os.path.join(ROOT, "oracle", "_ref", "hc_e2e_b200_batched")
GL_EXE = os.path.join(ROOT, "oracle", "_ref", "hc_e2e_b200_gl")
GOLDEN = os.path.join(ROOT, "tests", "golden", "chrm_like.ref.vcf")
needs = pytest.mark.skipif(not (os.path.exists(REF_EXE) and os.path.exists(B200_EXE)),
                           reason="oracle/_ref/hc_e2e_* not built (needs /root/reference at build time)")


def _run(exe, prefix, out, extra=()):
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-I", prefix + ".sam", "-R", prefix + ".fa", "-O", out, *extra], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return time.perf_counter() - t0, r.stderr.strip().splitlines()[-1]


@pytest.fixture(scope="module")
def data(tmp_path_factory, pkg):
    prefix = str(tmp_path_factory.mktemp("chrm") / "chrm_like")
    truth = pkg.synth.chrm_like(prefix)
    return prefix, truth


@needs
def test_reference_engine_end_to_end_matches_golden(data, tmp_path):
    prefix, truth = data
    out = str(tmp_path / "ref.vcf")
    _run(REF_EXE, prefix, out)
    got = open(out).read()
    assert got == open(GOLDEN).read()
    called = {int(l.split("\t")[1]) for l in got.splitlines() if not l.startswith("#")}
    snps = [p for p, kind, _, _ in truth if kind == "snp"]
    assert sum((p + 1) in called for p in snps) >= 0.9 * len(snps)      # the pipeline really calls the truth


@pytest.mark.gpu
@needs
def test_b200_engine_end_to_end_vcf_is_bit_identical(data, tmp_path):
    prefix, _ = data
    ref_out, b200_out = str(tmp_path / "ref.vcf"), str(tmp_path / "b200.vcf")
    t_ref, log_ref = _run(REF_EXE, prefix, ref_out)
    t_b200, log_b200 = _run(B200_EXE, prefix, b200_out)
    print(f"\nchrM-like e2e wall: reference engine {t_ref:.2f} s ({log_ref}) | B200 engine {t_b200:.2f} s ({log_b200})")
    assert open(b200_out, "rb").read() == open(ref_out, "rb").read()
    assert open(b200_out).read() == open(GOLDEN).read()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BATCHED_EXE), reason="oracle/_ref/hc_e2e_b200_batched not built")
@pytest.mark.parametrize("threads", [1, 8])
def test_batched_driver_vcf_is_bit_identical(data, tmp_path, threads):
    """windows assembled on `threads` host threads, all regions of the contig scored in a few
    asynchronous cross-window batches (hc::B200RegionBatcher), genotyped in window order."""
    prefix, _ = data
    out = str(tmp_path / "batched.vcf")
    t, log = _run(BATCHED_EXE, prefix, out, extra=("-T", str(threads)))
    print(f"\nchrM-like e2e wall, batched driver, {threads} assembly thread(s): {t:.2f} s ({log})")
    assert open(out).read() == open(GOLDEN).read()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(GL_EXE), reason="oracle/_ref/hc_e2e_b200_gl not built")
def test_device_genotype_reduction_vcf_is_bit_identical(data, tmp_path):
    """SURVEY 8f-3 end to end: the reference's driver with Genetyper::assign_genotype_likelihoods cut in two around
    the likelihoods -- sites planned on the host (the reference's own helpers), cap / filter / marginalisation /
    genotype likelihoods computed ON THE DEVICE (phmm_submit_gl), genotype quality and the call on the host.
    The reads x haplotypes matrix never comes back; the VCF is byte-identical to the reference's."""
    prefix, _ = data
    out = str(tmp_path / "gl.vcf")
    t, log = _run(GL_EXE, prefix, out, extra=("-T", "8"))
    print(f"\nchrM-like e2e wall, batched driver with device-side genotype likelihoods: {t:.2f} s ({log})")
    assert open(out).read() == open(GOLDEN).read()
