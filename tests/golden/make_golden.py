"""Generates the golden fixtures of tests/golden/ by running the REFERENCE's own PairHMM code.

Run in the build container only (it needs oracle/_ref/libref_pairhmm.so, which oracle/Makefile
compiles from /root/reference in place).  The fixtures are committed; the GPU box never runs this.

  kat_appendix_a.json       the known-answer vectors of SURVEY.md Appendix A.1 (kernel level) and
                            A.2 (hc::IntelPairHMM::compute_likelihoods), re-derived here from the
                            compiled reference and cross-checked against the values quoted in SURVEY.md
  ref_random_pairs.json     seeded random pairs (N bases, lower case, per-base gap penalties, ragged
                            lengths) -> raw f32 bits, raw f64 bits, rescue flag, final log10
  ref_region_filter.json    regions through compute_likelihoods: kept reads and capped matrix
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _oracle import load_ref  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

ref = load_ref()
assert ref is not None, "oracle/_ref/libref_pairhmm.so missing: make -C oracle"
pkg = load_package()

f32bits = lambda x: struct.unpack("<I", struct.pack("<f", x))[0]
f64bits = lambda x: struct.unpack("<Q", struct.pack("<d", x))[0]
H40 = "ACGTACGTTTGACCAGTACGATCGATCGGATCGATTTAGC"


def one(read, qual, hap):
    n = len(read)
    i = b"I" * n; c = b"+" * n
    f = ref.forward(read.encode(), qual.encode(), i, i, c, hap.encode(), "f32")
    d = ref.forward(read.encode(), qual.encode(), i, i, c, hap.encode(), "f64")
    return {"read": read, "qual": qual, "hap": hap, "f32_bits": f32bits(f), "f64_bits": f64bits(d),
            "rescue": int(f < np.float32(1e-28))}


# ---- Appendix A.1 : inputs as listed there, expected bits as quoted there ----
survey_a1 = [
    ("1 exact", "ACGTTTGACCAGTACGATC", "F" * 19, H40, 0x78ccc9e3, 0x7f59993f6268319b, 0),
    ("2 one mismatch", "ACGTTTGACCAGAACGATC", "F" * 19, H40, 0x6c650d62, 0x7dcca1af44f27685, 0),
    ("3 mismatch, low q", "ACGTTTGACCAGAACGATC", "5" * 19, H40, 0x6f335b36, 0x7e266b67ba886609, 0),
    ("4 N in read", "ACGTTTGACCAGNACGATC", "F" * 19, H40, 0x78ccc9e3, 0x7f59993f626831bf, 0),
    ("5 1-bp deletion", "ACGTTTGACCAGACGATC", "F" * 18, H40, 0x6cac3044, 0x7dd5860a9c06207f, 0),
    ("6 1-bp insertion", "ACGTTTGACCAGTTACGATC", "F" * 20, H40, 0x6d2c3041, 0x7de5860a47ae1bbe, 0),
    ("7 9 rows", "ACGTTTGAC", "F" * 9, H40, 0x78ccca0c, 0x7f599942d3e6099a, 0),
    ("8 8 rows", "ACGTTTGA", "F" * 8, H40, 0x78ccca0f, 0x7f599943145f30c4, 0),
    ("9 all mismatch", "T" * 20, "I" * 20, H40, 0x00000000, 0x6fb672c071566b58, 1),
    ("10 lower-case read", "acgtttgaccagtacgatc", "F" * 19, H40, 0x00000000, 0x6ede9dfed0ff4497, 1),
    ("11 N in haplotype", "ACGTTTGACCAGTACGATC", "F" * 19, H40[:14] + "N" + H40[15:], 0x78ccc9e3, 0x7f59993f626831d1, 0),
]
kats = []
for name, read, qual, hap, b32, b64, resc in survey_a1:
    got = one(read, qual, hap)
    assert (got["f32_bits"], got["f64_bits"], got["rescue"]) == (b32, b64, resc), (name, got)
    got["name"] = name
    kats.append(got)

# ---- Appendix A.2 : call surface ----
h1 = H40[:16] + "A" + H40[17:]
reads = ["ACGTTTGACCAGTACGATC", "ACGTTTGACCAGAACGATC", "T" * 20]
quals = ["F" * 19, "5" * 19, "I" * 20]


def call_surface(reads, quals, haps):
    ro = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int32)
    ho = np.concatenate([[0], np.cumsum([len(h) for h in haps])]).astype(np.int32)
    rb = np.frombuffer("".join(reads).encode(), np.uint8).copy()
    rq = np.frombuffer("".join(quals).encode(), np.uint8).copy()
    hb = np.frombuffer("".join(haps).encode(), np.uint8).copy()
    lik = np.zeros(len(reads) * len(haps)); keep = np.zeros(len(reads), np.uint8)
    n = ref.lib.ref_compute_likelihoods(len(reads), ro, rb, rq, len(haps), ho, hb, lik, keep)
    return keep.tolist(), lik[: n * len(haps)].reshape(n, len(haps)).tolist()


keep, lik = call_surface(reads, quals, [H40, h1])
assert keep == [1, 1, 0]
assert abs(lik[0][0] + 1.602085114) < 1e-8 and abs(lik[0][1] + 6.102085114) < 1e-8
assert abs(lik[1][0] + 6.102123260) < 1e-8 and abs(lik[1][1] + 1.602123260) < 1e-8
a2 = {"haps": [H40, h1], "reads": reads, "quals": quals, "keep": keep, "lik": lik}
l32, l64 = ref.log10_init()
json.dump({"source": "reference compiled from /root/reference (oracle/_ref), cross-checked with SURVEY.md Appendix A",
           "log10_init_f32": l32, "log10_init_f64": l64, "a1": kats, "a2": a2},
          open(os.path.join(HERE, "kat_appendix_a.json"), "w"), indent=1)

# ---- seeded random pairs ----
pairs = []
for seed, kw in ((11, dict(general_gaps=True, n_frac=0.03)), (12, dict(general_gaps=False, n_frac=0.0)),
                 (13, dict(general_gaps=True, n_frac=0.05, lower_frac=0.05, max_read_len=200, max_hap_len=400)),
                 (14, dict(general_gaps=False, max_read_len=255, max_hap_len=300, n_regions=2))):
    b = pkg.synth.random_small(seed, **kw)
    out = ref.batch(b)
    pairs.append({"seed": seed, "kw": kw, "n_pairs": b.n_pairs,
                  "raw32_bits": out["raw32"].view(np.uint32).tolist(),
                  "raw64_bits": out["raw64"].view(np.uint64).tolist(),
                  "rescued": out["rescued"].tolist(),
                  "log10_bits": out["log10"].view(np.uint64).tolist()})
json.dump({"source": "ref_batch() of oracle/_ref over phmm_b200.synth.random_small(seed, **kw)", "batches": pairs},
          open(os.path.join(HERE, "ref_random_pairs.json"), "w"))

# ---- regions through compute_likelihoods (cap + poorly-modelled filter) ----
regions = []
rng = np.random.default_rng(21)
for t in range(6):
    b = pkg.synth.random_small(100 + t, n_regions=1, max_reads=12, max_haps=6, general_gaps=False)
    rd = [bytes(b.read_bases[b.read_off[i]:b.read_off[i + 1]]).decode() for i in range(b.n_reads)]
    ql = [bytes(b.read_q[b.read_off[i]:b.read_off[i + 1]]).decode() for i in range(b.n_reads)]
    hp = [bytes(b.hap_bases[b.hap_off[i]:b.hap_off[i + 1]]).decode() for i in range(b.n_haps)]
    keep, lik = call_surface(rd, ql, hp)
    regions.append({"reads": rd, "quals": ql, "haps": hp, "keep": keep,
                    "lik_bits": np.array(lik, np.float64).reshape(-1).view(np.uint64).tolist()})
json.dump({"source": "hc::IntelPairHMM::compute_likelihoods of the compiled reference", "regions": regions},
          open(os.path.join(HERE, "ref_region_filter.json"), "w"))
print("golden fixtures written:", [f for f in os.listdir(HERE) if f.endswith(".json")])
