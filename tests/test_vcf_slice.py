"""VCF-level parity: the reference's own SW aligner, Genotyper and Variant::print
around either likelihood engine (oracle/vcf_slice.cpp -> oracle/_ref/vcf_slice).

  CPU : the committed golden VCF is what the REFERENCE engine produces (regenerated here when the
        harness binary is present);
  GPU : hc::B200PairHMM (include/b200_pairhmm.hpp over the C ABI) yields a BYTE-IDENTICAL VCF, the same
        surviving reads and likelihoods within 1e-4.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "oracle", "_ref", "vcf_slice")
needs_harness = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/vcf_slice not built (needs /root/reference at build time)")


def _read_lik(path):
    out = []
    for line in open(path):
        p = line.split()
        if p[0] == "REGION":
            out.append({"kept": int(p[3]), "rows": {}})
        else:
            out[-1]["rows"][p[0]] = np.array([float(x) for x in p[1:]])
    return out


def test_golden_vcf_is_well_formed():
    lines = open(os.path.join(G, "vcf_regions.ref.vcf")).read().splitlines()
    assert len(lines) == 13
    for ln in lines:
        f = ln.split("\t")
        assert f[0] == "chrS" and f[8] == "GT:GQ" and f[9].split(":")[0] in ("0/1", "1/1", "1/2")


@needs_harness
def test_reference_engine_reproduces_golden(tmp_path):
    dump = tmp_path / "lik.txt"
    r = subprocess.run([EXE, "--engine", "ref", "--dump", str(dump), os.path.join(G, "vcf_regions.txt")],
                       capture_output=True, text=True, check=True)
    assert r.stdout == open(os.path.join(G, "vcf_regions.ref.vcf")).read()
    assert open(dump).read() == open(os.path.join(G, "vcf_regions.ref.lik")).read()


@pytest.mark.gpu
@needs_harness
def test_b200_engine_gives_byte_identical_vcf(tmp_path):
    dump = tmp_path / "lik_b200.txt"
    r = subprocess.run([EXE, "--engine", "b200", "--dump", str(dump), os.path.join(G, "vcf_regions.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(G, "vcf_regions.ref.vcf")).read()          # byte for byte
    got, want = _read_lik(dump), _read_lik(os.path.join(G, "vcf_regions.ref.lik"))
    assert len(got) == len(want) == 24
    worst = 0.0
    for a, b in zip(got, want):
        assert a["kept"] == b["kept"] and a["rows"].keys() == b["rows"].keys()       # same reads survive
        for k in a["rows"]:
            worst = max(worst, float(np.abs(a["rows"][k] - b["rows"][k]).max()))
    assert worst <= 1e-4, worst
