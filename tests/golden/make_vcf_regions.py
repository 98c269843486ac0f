"""Writes tests/golden/vcf_regions.txt (synthetic active regions for oracle/vcf_slice.cpp) and, by running
the REFERENCE engine on them (oracle/_ref/vcf_slice --engine ref, CPU), the golden outputs
vcf_regions.ref.vcf and vcf_regions.ref.lik.  Build-container only; the outputs are committed.

Regions follow the reference's windowing (245 bp windows padded by 85 bp, haplotypecaller.hpp:112-113,
:126-128): a random reference window, 2..6 candidate haplotypes (the reference path plus SNP /
insertion / deletion haplotypes inside the origin window), and ~30x of 100..150 bp diploid reads with
1 % substitutions, Q in [20,40], clipped to the padded window."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ACGT = np.frombuffer(b"ACGT", np.uint8)
rng = np.random.default_rng(4242)
lines = []
contig, pos = "chrS", 0
for w in range(24):
    pb, pe = (0, 330) if w == 0 else (pos - 85, pos + 245 + 85)
    ob, oe = pos, pos + 245
    ref = ACGT[rng.integers(0, 4, pe - pb)].tobytes().decode()
    n_alt = int(rng.integers(1, 6))
    haps = [ref]
    for _ in range(n_alt):
        p = int(rng.integers(ob - pb + 10, oe - pb - 10))
        kind = rng.choice(["snp", "snp", "ins", "del"])
        if kind == "snp":
            alt = "ACGT"[("ACGT".index(ref[p]) + int(rng.integers(1, 4))) % 4]
            h = ref[:p] + alt + ref[p + 1:]
        elif kind == "ins":
            ins = ACGT[rng.integers(0, 4, int(rng.integers(1, 5)))].tobytes().decode()
            h = ref[:p] + ins + ref[p:]
        else:
            d = int(rng.integers(1, 5))
            h = ref[:p] + ref[p + d:]
        if h not in haps:
            haps.append(h)
    truth = [haps[0], haps[int(rng.integers(0, len(haps)))]] if rng.random() < 0.8 else [haps[-1], haps[-1]]
    lines.append(f"REGION {contig} {pb} {pe} {ob} {oe}")
    lines.append(f"REF {ref}")
    lines += [f"H {h}" for h in haps]
    n_reads = int(30 * (pe - pb) / 150)
    for _ in range(n_reads):
        t = truth[int(rng.integers(0, 2))]
        rl = int(rng.integers(100, 151))
        start = int(rng.integers(0, max(1, len(t) - rl)))
        seq = np.frombuffer(t[start:start + rl].encode(), np.uint8).copy()
        sub = rng.random(len(seq)) < 0.01
        seq[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
        qual = (33 + rng.integers(20, 41, len(seq))).astype(np.uint8)
        lines.append(f"R {pb + start + 1} {len(seq)}M {seq.tobytes().decode()} {qual.tobytes().decode()}")
    lines.append("END")
    pos += 245
path = os.path.join(HERE, "vcf_regions.txt")
open(path, "w").write("\n".join(lines) + "\n")
exe = os.path.join(ROOT, "oracle", "_ref", "vcf_slice")
vcf = subprocess.run([exe, "--engine", "ref", "--dump", os.path.join(HERE, "vcf_regions.ref.lik"), path],
                     check=True, capture_output=True, text=True)
open(os.path.join(HERE, "vcf_regions.ref.vcf"), "w").write(vcf.stdout)
print(vcf.stderr.strip(), "| variants:", vcf.stdout.count("\n"))
print(vcf.stdout[:600])
