"""Synthetic PairHMM workloads S2..S5 of SURVEY.md section 8(d) / BASELINE.json `configs`.

Bases are uniform over ACGT, quality bytes are ASCII Phred+33 exactly as the reference feeds them
to its kernel (no subtraction); gap-open/continuation default to the
reference's constant 'I' / 'I' / '+' strings (sam/sam.hpp:30-32), `general_gaps=True` draws
per-base i,d in '!'+[20,50) and c in '!'+[5,25) to exercise the per-row path.
Generators are numpy (PCG64) with fixed seeds so the GPU engine, the oracle and the compiled
reference all see byte-identical batches.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)
SEEDS = {"S2": 1002, "S3": 1003, "S4": 1004, "S5": 1005}


def _batch_cls():
    from . import Batch
    return Batch


def _gaps(rng, n, general):
    if not general:
        return {}
    return dict(read_i=(33 + rng.integers(20, 50, n)).astype(np.uint8),
                read_d=(33 + rng.integers(20, 50, n)).astype(np.uint8),
                read_c=(33 + rng.integers(5, 25, n)).astype(np.uint8))


def fixed_shape(n_regions, read_len, hap_len, n_reads, n_haps, seed, snps_per_hap=3, sub_rate=0.01,
                q_lo=20, q_hi=40, general_gaps=False):
    """S2/S3 family: per region one random backbone, haplotypes = backbone + SNPs, reads = noisy
    substrings of a random haplotype at a random offset."""
    rng = np.random.default_rng(seed)
    backbone = rng.integers(0, 4, (n_regions, 1, hap_len), dtype=np.int8)
    haps = np.repeat(backbone, n_haps, axis=1)
    pos = rng.integers(0, hap_len, (n_regions, n_haps, snps_per_hap))
    delta = rng.integers(1, 4, (n_regions, n_haps, snps_per_hap), dtype=np.int8)
    gi, hi = np.meshgrid(np.arange(n_regions), np.arange(n_haps), indexing="ij")
    for s in range(snps_per_hap):
        haps[gi, hi, pos[:, :, s]] = (haps[gi, hi, pos[:, :, s]] + delta[:, :, s]) % 4
    src = rng.integers(0, n_haps, (n_regions, n_reads))
    off = rng.integers(0, hap_len - read_len + 1, (n_regions, n_reads))
    idx = off[:, :, None] + np.arange(read_len)[None, None, :]
    reads = haps[np.arange(n_regions)[:, None, None], src[:, :, None], idx]
    sub = rng.random(reads.shape) < sub_rate
    reads = np.where(sub, (reads + rng.integers(1, 4, reads.shape, dtype=np.int8)) % 4, reads)
    quals = (33 + rng.integers(q_lo, q_hi + 1, reads.shape)).astype(np.uint8)
    n_r, n_h = n_regions * n_reads, n_regions * n_haps
    B = _batch_cls()
    return B(np.arange(n_regions + 1) * n_reads, np.arange(n_regions + 1) * n_haps,
             np.arange(n_r + 1) * read_len, ACGT[reads.reshape(-1)], quals.reshape(-1),
             np.arange(n_h + 1) * hap_len, ACGT[haps.reshape(-1)],
             **_gaps(rng, n_r * read_len, general_gaps))


def s2(n_regions=256, general_gaps=False, seed=SEEDS["S2"]):
    """BASELINE config 2: 100 bp reads x 300 bp haplotypes, 64 reads x 8 haplotypes per region."""
    return fixed_shape(n_regions, 100, 300, 64, 8, seed, general_gaps=general_gaps)


def s3(n_regions=64, general_gaps=False, seed=SEEDS["S3"]):
    """BASELINE config 3 (the metric's workload): 150 bp x 500 bp, 256 reads x 16 haplotypes."""
    return fixed_shape(n_regions, 150, 500, 256, 16, seed, general_gaps=general_gaps)


def s4(n_regions=16, n_reads=128, n_haps=16, general_gaps=False, seed=SEEDS["S4"],
       read_lo=150, read_hi=250, hap_lo=600, hap_hi=1000, tail=60):
    """BASELINE config 4: long ragged pairs; the last `tail` read bases are random with Q in [2,10],
    which drives the FP32 result under 1e-28 and forces the FP64 redo (intel_pairhmm.hpp:137)."""
    rng = np.random.default_rng(seed)
    regions = []
    for _ in range(n_regions):
        hl = int(rng.integers(hap_lo, hap_hi + 1))
        backbone = rng.integers(0, 4, hl, dtype=np.int8)
        haps = []
        for _h in range(n_haps):
            h = backbone.copy()
            p = rng.integers(0, hl, 3)
            h[p] = (h[p] + rng.integers(1, 4, 3)) % 4
            haps.append(ACGT[h])
        reads, quals = [], []
        for _r in range(n_reads):
            rl = int(rng.integers(read_lo, min(read_hi, hl) + 1))
            src = haps[int(rng.integers(0, n_haps))]
            o = int(rng.integers(0, hl - rl + 1))
            r = src[o:o + rl].copy()
            sub = rng.random(rl) < 0.01
            r[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
            q = (33 + rng.integers(20, 41, rl)).astype(np.uint8)
            t = min(tail, rl)
            r[rl - t:] = ACGT[rng.integers(0, 4, t)]
            q[rl - t:] = (33 + rng.integers(2, 11, t)).astype(np.uint8)
            reads.append(r); quals.append(q)
        if general_gaps:
            gi = [(33 + rng.integers(20, 50, len(r))).astype(np.uint8) for r in reads]
            gd = [(33 + rng.integers(20, 50, len(r))).astype(np.uint8) for r in reads]
            gc = [(33 + rng.integers(5, 25, len(r))).astype(np.uint8) for r in reads]
            regions.append((reads, quals, haps, gi, gd, gc))
        else:
            regions.append((reads, quals, haps))
    return _batch_cls().from_regions(regions)


def s5_stream(n_windows, seed=SEEDS["S5"], coverage=30, read_len=150, window=245, pad=85,
              snp_rate=1e-3, indel_rate=1e-4, windows_per_batch=256):
    """BASELINE config 5: a stream of active regions cut from a synthetic reference in the
    reference's own windowing (245 bp windows, 85 bp padding, haplotypecaller.hpp:112-113,126-128).
    The assembler is out of scope, so per-window haplotypes are synthesised: the padded reference
    window plus combinations of the truth variants falling in it (2..16 haplotypes).  Reads are drawn
    from one of two truth haplotypes (diploid), clipped to the padded window like
    ReadClipper::hard_clip_to_interval does, 1 % substitutions, Q in [20,40].
    Yields Batch objects of `windows_per_batch` windows; the reference sequence is generated
    window by window so the stream has no length limit."""
    rng = np.random.default_rng(seed)
    B = _batch_cls()
    span = window + 2 * pad
    regions = []
    for _w in range(n_windows):
        ref = rng.integers(0, 4, span, dtype=np.int8)
        n_snp = rng.binomial(span, snp_rate)
        n_indel = rng.binomial(span, indel_rate)
        events = []
        for p in sorted(rng.choice(span - 20, size=min(n_snp + n_indel, 6), replace=False) + 10):
            kind = "snp" if rng.random() < snp_rate / (snp_rate + indel_rate) else ("ins" if rng.random() < 0.5 else "del")
            events.append((int(p), kind, int(rng.integers(1, 4)), int(rng.integers(1, 6))))

        def apply(mask):
            out, last = [], 0
            for k, (p, kind, d, ln) in enumerate(events):
                if not (mask >> k) & 1:
                    continue
                out.append(ref[last:p])
                if kind == "snp":
                    out.append(np.array([(ref[p] + d) % 4], np.int8)); last = p + 1
                elif kind == "ins":
                    out.append(ref[p:p + 1]); out.append(((ref[p] + d + np.arange(ln)) % 4).astype(np.int8)); last = p + 1
                else:
                    last = min(span, p + ln)
            out.append(ref[last:])
            return np.concatenate(out)

        n_ev = len(events)
        masks = list(range(min(1 << n_ev, 16))) if n_ev else [0]
        if len(masks) < 2:
            masks = [0, 0]          # the driver skips regions with <= 1 haplotype (haplotypecaller.hpp:101);
                                    # keep two so the stream still exercises the engine
        haps = [ACGT[apply(m)] for m in masks]
        truth = [haps[0], haps[int(rng.integers(0, len(haps)))]]
        n_reads = int(rng.poisson(coverage * span / read_len))
        n_reads = max(1, min(n_reads, span))
        reads, quals = [], []
        for _r in range(n_reads):
            t = truth[int(rng.integers(0, 2))]
            start = int(rng.integers(-read_len + 10, len(t) - 10))
            a, b = max(0, start), min(len(t), start + read_len)
            r = t[a:b].copy()
            sub = rng.random(len(r)) < 0.01
            r[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
            reads.append(r); quals.append((33 + rng.integers(20, 41, len(r))).astype(np.uint8))
        regions.append((reads, quals, haps))
        if len(regions) == windows_per_batch:
            yield B.from_regions(regions)
            regions = []
    if regions:
        yield B.from_regions(regions)


def _s5_chunk(args):
    n, seed = args
    return next(s5_stream(n, seed=seed, windows_per_batch=n))


def s5_batch(n_windows, seed=SEEDS["S5"], procs=None, chunk=256):
    """One S5 batch of `n_windows` windows, generated in `chunk`-window pieces (piece k from seed + k, so the
    result does not depend on the process count) on a process pool: the per-read python loop of s5_stream
    makes ~350 windows/s on one core, far too slow for 8-GPU sized batches."""
    import os
    B = _batch_cls()
    jobs = [(min(chunk, n_windows - o), seed + 7919 * k) for k, o in enumerate(range(0, n_windows, chunk))]
    procs = procs or min(len(jobs), os.cpu_count() or 1)
    if procs <= 1 or len(jobs) == 1:
        return B.concat([_s5_chunk(j) for j in jobs])
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs) as pool:
        return B.concat(pool.map(_s5_chunk, jobs))


def sw_pairs(n_regions, n_haps=16, seed=11, window=415):
    """Smith-Waterman workload: every haplotype of a padded window aligned to that window, as
    assembler/graph_wrapper.hpp:232-239 does; haplotypes = the window with 1-4 SNPs / small indels each.
    Returns [(window bytes, haplotype bytes)]."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_regions):
        ref = ACGT[rng.integers(0, 4, window)]
        for _h in range(n_haps):
            alt = list(ref)
            for _e in range(int(rng.integers(1, 5))):
                k = int(rng.integers(10, len(alt) - 10)); t = int(rng.integers(0, 3))
                if t == 0:
                    alt[k] = int(ACGT[rng.integers(0, 4)])
                elif t == 1:
                    del alt[k:k + int(rng.integers(1, 8))]
                else:
                    alt[k:k] = [int(x) for x in ACGT[rng.integers(0, 4, int(rng.integers(1, 8)))]]
            out.append((ref.tobytes(), np.array(alt, np.uint8).tobytes()))
    return out


def random_small(seed, n_regions=3, max_reads=9, max_haps=5, max_read_len=70, max_hap_len=120,
                 general_gaps=True, n_frac=0.03, lower_frac=0.0):
    """Ragged little regions with N bases and arbitrary qualities: parity-test fodder."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGTN", np.uint8)
    p = np.array([1 - n_frac] * 4 + [4 * n_frac]) / 4
    p = p / p.sum()
    regions = []
    for _ in range(n_regions):
        nh = int(rng.integers(1, max_haps + 1)); nr = int(rng.integers(1, max_reads + 1))
        haps = [alpha[rng.choice(5, int(rng.integers(1, max_hap_len + 1)), p=p)] for _ in range(nh)]
        reads, quals, gi, gd, gc = [], [], [], [], []
        for _r in range(nr):
            rl = int(rng.integers(1, max_read_len + 1))
            h = haps[int(rng.integers(0, nh))]
            if rl <= len(h) and rng.random() < 0.7:
                o = int(rng.integers(0, len(h) - rl + 1)); r = h[o:o + rl].copy()
                m = rng.random(rl) < 0.04; r[m] = alpha[rng.integers(0, 4, int(m.sum()))]
            else:
                r = alpha[rng.choice(5, rl, p=p)]
            if lower_frac:
                m = rng.random(rl) < lower_frac; r = np.where(m, r | 0x20, r).astype(np.uint8)
            reads.append(r)
            quals.append((33 + rng.integers(2, 42, rl)).astype(np.uint8))
            gi.append((33 + rng.integers(20, 50, rl)).astype(np.uint8))
            gd.append((33 + rng.integers(20, 50, rl)).astype(np.uint8))
            gc.append((33 + rng.integers(5, 25, rl)).astype(np.uint8))
        regions.append((reads, quals, haps, gi, gd, gc) if general_gaps else (reads, quals, haps))
    return _batch_cls().from_regions(regions)


def chrm_like(prefix, length=16569, seed=1001, read_len=150, density=0.30, snp_every=350, indel_every=1400,
              contig="chrM"):
    """BASELINE config 1 stand-in (S1): the reference's own chrM files are not shipped, so this writes a
    synthetic chrM-like `prefix.fa` + `prefix.sam` for the end-to-end harness (oracle/hc_e2e.cpp):
    a random 16 569 bp contig, a truth set of SNPs and small indels (het or hom) carried by two phased
    haplotypes, and 150 bp reads with AT MOST ONE READ PER START POSITION (the reference picks one read
    per start with std::random_device, haplotypecaller.hpp:44-50; one candidate makes the draw forced).
    MAPQ 60, RNEXT '=', proper CIGARs for reads spanning indels, 0.5 % substitutions, Q in [25,40].
    Returns the truth list [(pos0, kind, alt, genotype)]."""
    rng = np.random.default_rng(seed)
    ref = ACGT[rng.integers(0, 4, length)]
    truth = []                       # spaced so that the variants do not interact
    pos = 200
    while pos < length - 400:
        kind = "snp"
        if rng.random() < snp_every / indel_every:
            kind = "ins" if rng.random() < 0.5 else "del"
        gt = "hom" if rng.random() < 0.3 else "het"
        if kind == "snp":
            alt = ACGT[(int(np.where(ACGT == ref[pos])[0][0]) + int(rng.integers(1, 4))) % 4]
            truth.append((pos, "snp", bytes([alt]), gt))
        elif kind == "ins":
            truth.append((pos, "ins", ACGT[rng.integers(0, 4, int(rng.integers(1, 4)))].tobytes(), gt))
        else:
            truth.append((pos, "del", int(rng.integers(1, 4)), gt))
        pos += int(rng.integers(snp_every // 2, snp_every * 3 // 2))
    with open(prefix + ".fa", "w") as f:
        f.write(f">{contig}\n")
        s = ref.tobytes().decode()
        for i in range(0, length, 60):
            f.write(s[i:i + 60] + "\n")

    def make_read(start, hap):
        """Walk the reference from `start`, applying the variants this haplotype carries."""
        seq, cigar, p, n_m = [], [], start, 0
        vi = {v[0]: v for v in truth if v[3] == "hom" or hap == 1}
        while len(seq) < read_len and p < length:
            v = vi.get(p)
            if v is None or v[1] == "snp":
                seq.append(v[2][0] if v is not None else ref[p]); n_m += 1; p += 1
            elif v[1] == "ins":                       # anchor base, then the inserted bases
                seq.append(ref[p]); n_m += 1; p += 1
                ins = list(v[2])[: read_len - len(seq)]
                if ins:
                    cigar.append(f"{n_m}M{len(ins)}I"); n_m = 0; seq += ins
            else:                                     # anchor base, then skip the deleted bases
                seq.append(ref[p]); n_m += 1; p += 1
                if len(seq) < read_len and p + v[2] < length:
                    cigar.append(f"{n_m}M{v[2]}D"); n_m = 0; p += v[2]
        if n_m:
            cigar.append(f"{n_m}M")
        if cigar and cigar[-1].endswith(("I", "D")):   # never end on an indel
            return None
        return np.array(seq, np.uint8), "".join(cigar)

    n = 0
    with open(prefix + ".sam", "w") as f:
        f.write(f"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:{contig}\tLN:{length}\n")
        for start in range(0, length - read_len - 8):
            if rng.random() > density:
                continue
            r = make_read(start, int(rng.integers(0, 2)))
            if r is None:
                continue
            seq, cigar = r
            sub = rng.random(len(seq)) < 0.005
            seq = seq.copy(); seq[sub] = ACGT[rng.integers(0, 4, int(sub.sum()))]
            qual = (33 + rng.integers(25, 41, len(seq))).astype(np.uint8)
            f.write(f"r{n}\t0\t{contig}\t{start + 1}\t60\t{cigar}\t=\t{start + 1}\t0\t{seq.tobytes().decode()}\t{qual.tobytes().decode()}\n")
            n += 1
    return truth
