#!/usr/bin/env python
"""bench.py -- PairHMM forward throughput (GCUPS) on B200, the metric of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W [--workload s2|s3|s4|s5]     (our arm; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W [--workload ..]   (the reference's CPU PairHMM)

Workloads (config.workload) are the synthetic shapes of SURVEY.md section 8(d) / BASELINE.json configs[1..4];
the default, S3, is the one the metric is quoted on: 150 bp reads x 500 bp haplotypes, 256 reads x 16
haplotypes per active region, 128 regions per GPU per step.  A step = one pass of the hot path (FP32 forward
kernels + FP64 redo kernels) over one batch.  GCUPS = sum over pairs of read_len x hap_len / seconds / 1e9,
each pair once.

  value      inputs already resident in HBM (phmm_stage), kernels only, CUDA events on the launching
             streams (inside the C library), max over ranks, summed over ranks' cells.
  parity     every device-resident batch the timed region ran is fetched afterwards and >= 2 whole regions
             of each are compared with the CPU oracle (1e-4 FP32 path / 1e-9 FP64-rescued, rescue decisions
             equal); the checker runs in a SUBPROCESS (this process never maps anything under oracle/) and
             a failure makes the run exit non-zero.
  e2e        the same batches through the C ABI with HOST buffers (phmm_submit / phmm_wait, 4 batches in
             flight): H2D of the step's inputs from page-locked host memory + kernels + D2H + host log10,
             wall clock.  `e2e_pageable`: the same with ordinary (pageable) numpy arrays, which the engine
             first copies into its own pinned staging.
  roofline   FP32 CUDA-core issue roofline of SURVEY.md section 8(d): SMs x 128 lanes x f_SM / 8
             FP32-pipe instructions per cell (NOT HBM: 7.6e-4 B/cell); `peak` uses the max SM clock
             of MEASURED_PEAKS.json, `peak_at_clock` the median clock sampled during the run.
  in_process (N > 1) after the per-rank section rank 0 alone drives ONE engine with n_devices = N -- the
             multi-GPU scheduler north_star names (one host process, a worker per device, regions sharded
             by cell count, no collective): S3 with N x 128 regions per step and the ragged S5 window
             stream, pageable and page-locked inputs, with a bit-for-bit comparison against the 1-device
             engine and an oracle check of 3 regions per device share.
  e2e_chrm   (N = 1) BASELINE.json's second metric: wall time of the reference's whole driver on a synthetic
             chrM-like contig around its own CPU engine and around this engine (batched driver), best of 3.
  cpu_baseline  the reference's own AVX PairHMM (oracle/_ref, kind "reference") -- or the oracle port
             when that library is absent -- on this box's host cores, bounded sample, rank 0, N=1,
             in a subprocess.

L2 rule: the timed steps rotate over NBATCH distinct device-resident batches whose inputs + outputs
exceed the 126 MB L2 (config.l2_policy).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pairhmm_gcups"
UNIT = "GCUPS"
SM_LANES = 128          # FP32 lanes per SM
INSTR_PER_CELL = 8      # FP32-pipe instructions per cell with FMA: the roofline's definition of a cell
INSTR_PER_CELL_SCALED = 5   # what the default engine's scaled, pMM-folded recurrence executes per cell (phmm_kernels.cuh: MODE 3, FOLD)
DEPTH = 4               # batches in flight on the e2e path


# ---- workloads ---------------------------------------------------------------------------------
class Workload:
    def __init__(self, key, label, regions, seed, make, unit="regions"):
        self.key, self.label, self.regions, self.seed, self._make, self.unit = key, label, regions, seed, make, unit

    def make(self, pkg, n, seed):
        return self._make(pkg.synth, n, seed)

    def name(self, n):
        return f"{self.label}, {n} {self.unit} per GPU per step, seed {self.seed}+"


WORKLOADS = {
    "s2": Workload("s2", "S2 (BASELINE configs[1]): synthetic PairHMM batch, 100bp reads x 300bp haplotypes, "
                   "64 reads x 8 haplotypes per region", 1024, 1002, lambda S, n, seed: S.s2(n, seed=seed)),
    "s3": Workload("s3", "S3 (BASELINE configs[2]): synthetic PairHMM batch, 150bp reads x 500bp haplotypes, "
                   "256 reads x 16 haplotypes per region", 128, 1003, lambda S, n, seed: S.s3(n, seed=seed)),
    "s3g": Workload("s3g", "S3 with per-base gap penalties (general mode of the ABI): 150bp x 500bp, 256 x 16 per region",
                    128, 1003, lambda S, n, seed: S.s3(n, general_gaps=True, seed=seed)),
    "s4": Workload("s4", "S4 (BASELINE configs[3]): long-pair mix, 150-250bp reads x 600-1000bp haplotypes with low-quality "
                   "tails (every pair redone in FP64), 128 reads x 16 haplotypes per region", 64, 1004,
                   lambda S, n, seed: S.s4(n, seed=seed)),
    "s5": Workload("s5", "S5 (BASELINE configs[4]): 30x active-region stream in the reference's 245+-85 windows, "
                   "2-16 synthesised haplotypes per window, reads clipped to the window", 1024, 1005,
                   lambda S, n, seed: S.s5_batch(n, seed=seed), unit="windows"),
}


# ---- distributed helpers (also exercised on CPU/gloo by tests/test_sharding_gloo.py) ------------
def _dist():
    import torch.distributed as dist
    return dist


def max_over_ranks(x, world):
    """MAX of a python float over ranks (timing rule: a multi-GPU number is the slowest rank's)."""
    if world == 1:
        return float(x)
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    if world == 1:
        return float(x)
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_results(local, offset, total, rank, world):
    """Host-side gather of per-rank result shards into the whole-batch vector on rank 0 (no collective
    sits on the compute path: this runs after the ranks have finished)."""
    if world == 1:
        return np.asarray(local)
    dist = _dist()
    parts = [None] * world if rank == 0 else None
    dist.gather_object((int(offset), np.asarray(local)), parts, dst=0)
    if rank != 0:
        return None
    out = np.empty(total, np.float64)
    for off, arr in parts:
        out[off:off + arr.size] = arr
    return out


# ---- clocks during the timed region -------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._live = threading.Event()
        self._live.set()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            if not self._live.is_set():                             # between timed regions: do not sample idle clocks
                self._live.wait(0.05)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def pause(self):
        self._live.clear()

    def resume(self):
        self._live.set()

    def __exit__(self, *a):
        self._stop.set()
        self._live.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.samples)}


# ---- CPU checkers: only ever inside a helper subprocess or the --impl reference arm --------------
def cpu_checker():
    """(checker, kind): the compiled reference if oracle/_ref travelled here, else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import load_oracle, load_ref
    ref = None
    try:
        ref = load_ref()
    except OSError:
        ref = None
    return (ref, "reference") if ref is not None else (load_oracle(), "port")


def cpu_gcups(pkg, wl, checker, n_regions, threads, seed):
    b = wl.make(pkg, n_regions, seed)
    t0 = time.perf_counter()
    checker.batch(b, threads=threads)
    dt = time.perf_counter() - t0
    return b.n_cells / dt / 1e9, dt, b


def helper_cpu_baseline(wl_key, budget_s=12.0):
    """--_helper cpu_baseline: prints the cpu_baseline object (run in a subprocess by the GPU arm)."""
    from __graft_entry__ import load_package
    pkg = load_package()
    wl = WORKLOADS[wl_key]
    checker, kind = cpu_checker()
    cores = os.cpu_count() or 1
    unit_n = max(1, wl.regions // 128)                              # smallest sample: ~1/128 of a step
    g1, _, _ = cpu_gcups(pkg, wl, checker, unit_n, 1, wl.seed)      # as shipped: OpenMP compiled out
    gp, _, bp = cpu_gcups(pkg, wl, checker, 2 * unit_n, cores, wl.seed)
    cells_per_region = bp.n_cells / bp.n_regions
    n = int(max(2 * unit_n, min(8 * wl.regions, budget_s * gp * 1e9 / cells_per_region)))
    gN, dtN, bN = cpu_gcups(pkg, wl, checker, n, cores, wl.seed)
    print(json.dumps({"value": round(gN, 3), "unit": UNIT, "cores": cores, "kind": kind,
                      "sample": f"{n} {wl.key.upper()} {wl.unit} ({bN.n_pairs} pairs, {bN.n_cells:.3e} cells) in {dtN:.1f} s with "
                                f"{cores} threads (omp dynamic over reads, intel_pairhmm.hpp:128-130)",
                      "as_shipped_1_thread": round(g1, 3)}))


def helper_sw_reference(n_pairs):
    """--_helper sw_reference: the reference's own AVX2 aligner (hc::IntelSWAligner::align, one thread, as it runs
    inside the assembler) on the first n_pairs of the bench's Smith-Waterman workload; prints GCUPS + the CIGARs' hash."""
    import ctypes as C
    import hashlib
    from __graft_entry__ import load_package
    pkg = load_package()
    path = os.path.join(ROOT, "oracle", "_ref", "libref_pairhmm.so")
    if not os.path.exists(path):
        print(json.dumps({"unavailable": "oracle/_ref not built"}))
        return
    lib = C.CDLL(path)
    lib.ref_sw_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    pairs = pkg.synth.sw_pairs((int(n_pairs) + 15) // 16)[:int(n_pairs)]
    buf = C.create_string_buffer(16384)
    h = hashlib.sha256()
    t0 = time.perf_counter()
    for r, a in pairs:
        off = lib.ref_sw_align(r, len(r), a, len(a), 200, -150, -260, -11, buf, 16384)
        h.update(f"{off}:{buf.value.decode()};".encode())
    dt = time.perf_counter() - t0
    cells = sum(len(r) * len(a) for r, a in pairs)
    print(json.dumps({"alignments": len(pairs), "gcups": round(cells / dt / 1e9, 3), "ms_per_alignment": round(1e3 * dt / len(pairs), 4),
                      "threads": 1, "sha256": h.hexdigest(), "what": "hc::IntelSWAligner::align (AVX2), compiled from the reference"}))


def sw_section(pkg, sms, sm_max_mhz):
    """the batched Smith-Waterman kernel behind phmm_sw_align: 16 384 alignments (1 024 windows of 415 bases x
    16 haplotypes), device time of the kernels and wall time of the whole call (strings in, CIGARs out)."""
    import hashlib
    pairs = pkg.synth.sw_pairs(1024)
    cells = sum(len(r) * len(a) for r, a in pairs)
    pkg.sw_align(pairs)
    best_call, best_k, got = 1e9, 1e9, None
    for _ in range(3):
        t0 = time.perf_counter()
        got, kms = pkg.sw_align(pairs)
        best_call, best_k = min(best_call, time.perf_counter() - t0), min(best_k, kms)
    # the C ABI alone (arrays in, arrays out): what a C++ caller pays, without python's string handling
    abi_s = pkg.sw_align_timed(pairs, repeats=3)
    ref = run_helper("sw_reference", "512")
    h = hashlib.sha256()
    for off, cg in got[:512]:
        h.update(f"{off}:{cg};".encode())
    ALU_OPS_PER_CELL = 14                 # MAIN_CODE of PairWiseSW.h:123-159: 5 adds, 3 max, 6 compare / select for the back-track bits
    peak = sms * 64 * sm_max_mhz * 1e6 / ALU_OPS_PER_CELL / 1e9
    return {"workload": "16 384 alignments: 1 024 windows of 415 bases x 16 haplotypes (1-4 SNPs / indels each), NEW_SW_PARAMETERS",
            "kernel_gcups": round(cells / best_k / 1e6, 1), "kernel_ms": round(best_k, 3),
            "call_gcups_c_abi": round(cells / abi_s / 1e9, 1), "call_ms_c_abi": round(1e3 * abi_s, 3),
            "call_over_kernel": round((best_k * 1e-3) / abi_s, 3),
            "call_gcups_python_strings": round(cells / best_call / 1e9, 1),
            "roofline": {"bound": "int32_alu", "peak": round(peak, 1), "frac": round(cells / best_k / 1e6 / peak, 3),
                         "peak_def": f"{sms} SMs x 64 INT32 lanes x {sm_max_mhz:.0f} MHz / {ALU_OPS_PER_CELL} ALU ops per cell"},
            "cpu_reference": ref, "identical_to_reference_on_sample": bool(ref.get("sha256") == h.hexdigest())}


def helper_parity(npz_path):
    """--_helper parity: the oracle over the region slices saved by the GPU arm; prints the parity object."""
    from __graft_entry__ import load_package
    pkg = load_package()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import load_oracle
    oracle = load_oracle()
    z = np.load(npz_path)
    n = int(z["n_slices"])
    e32 = e64 = 0.0
    mism = pairs = n_resc = flips = 0
    for k in range(n):
        g = lambda name: z[f"s{k}_{name}"]
        kw = {}
        if int(g("explicit")):
            kw = dict(read_i=g("read_i"), read_d=g("read_d"), read_c=g("read_c"))
        gi, gd, gc = (int(x) for x in g("gaps"))
        b = pkg.Batch(g("region_read_beg"), g("region_hap_beg"), g("read_off"), g("read_bases"), g("read_q"),
                      g("hap_off"), g("hap_bases"), gap_open_i=gi, gap_open_d=gd, gap_cont_c=gc, **kw)
        want = oracle.batch(b, threads=os.cpu_count() or 1)
        got, got_resc = g("log10"), g("rescued").astype(bool)
        resc = want["rescued"].astype(bool)
        diff = got_resc != resc
        if diff.any():
            # a rescue decision may legitimately differ only AT the threshold (raw FP32 sum within a few ulp of 1e-28f,
            # intel_pairhmm.hpp:137: the fast arithmetic rounds differently from the reference's) and then the two
            # values still agree to the FP32 tolerance; anything else is a mismatch
            thr = np.float32(1e-28).view(np.int32).astype(np.int64)
            ulps = np.abs(want["raw32"].view(np.int32).astype(np.int64) - thr)
            fine = diff & (ulps <= 16) & (np.abs(got - want["log10"]) <= 1e-4)
            flips += int(fine.sum())
            mism += int((diff & ~fine).sum())
        both32, both64 = ~resc & ~got_resc, resc & got_resc
        if both32.any():
            e32 = max(e32, float(np.abs(got[both32] - want["log10"][both32]).max()))
        if both64.any():
            d = np.abs(got[both64] - want["log10"][both64])
            d = d[np.isfinite(d) | (got[both64] != want["log10"][both64])]     # -inf == -inf is agreement
            if d.size:
                e64 = max(e64, float(np.nan_to_num(d, nan=np.inf, posinf=np.inf).max()))
        pairs += b.n_pairs
        n_resc += int(resc.sum())
    ok = mism == 0 and e32 <= 1e-4 and e64 <= 1e-9
    print(json.dumps({"max_abs_fp32": e32, "max_abs_fp64": e64, "rescue_mismatches": mism, "rescue_flips_at_threshold": flips, "regions_checked": n,
                      "pairs_checked": pairs, "rescued_pairs_checked": n_resc, "tolerance": "1e-4 fp32 / 1e-9 fp64-rescued",
                      "checker": "oracle/liboracle.so (subprocess)", "ok": bool(ok)}))


def run_helper(*argv, timeout=600):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--_helper", *argv], capture_output=True, text=True,
                       timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"helper {argv[0]} failed: {r.stderr[-800:]}")
    return json.loads(r.stdout.strip().splitlines()[-1])


class ParityCollector:
    """Region slices + the engine's results for them, checked by the oracle in a subprocess."""

    def __init__(self):
        self.items = {}
        self.n = 0

    def add(self, batch, log10, rescued, regions):
        ob = batch.region_out_beg
        for g in regions:
            sl = batch.slice_regions(g, g + 1)
            k = self.n
            self.n += 1
            for name in ("region_read_beg", "region_hap_beg", "read_off", "read_bases", "read_q", "hap_off", "hap_bases"):
                self.items[f"s{k}_{name}"] = getattr(sl, name)
            self.items[f"s{k}_explicit"] = np.int32(1 if sl.explicit_gaps else 0)
            if sl.explicit_gaps:
                for name in ("read_i", "read_d", "read_c"):
                    self.items[f"s{k}_{name}"] = getattr(sl, name)
            self.items[f"s{k}_gaps"] = np.array([sl.gap_open_i, sl.gap_open_d, sl.gap_cont_c], np.int32)
            self.items[f"s{k}_log10"] = np.asarray(log10[int(ob[g]):int(ob[g + 1])])
            self.items[f"s{k}_rescued"] = np.asarray(rescued[int(ob[g]):int(ob[g + 1])])

    def check(self):
        if not self.n:
            return None
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "parity.npz")
            np.savez(path, n_slices=np.int32(self.n), **self.items)
            return run_helper("parity", path)


def pick_regions(batch, k, seed):
    rng = np.random.default_rng(seed)
    live = np.nonzero(batch.reads_per_region * batch.haps_per_region > 0)[0]
    return sorted(int(g) for g in rng.choice(live, size=min(k, len(live)), replace=False))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from __graft_entry__ import load_package
    pkg = load_package()
    wl = WORKLOADS[args.workload]
    regions = args.regions or wl.regions
    checker, kind = cpu_checker()
    cores = os.cpu_count() or 1
    unit_n = max(1, wl.regions // 128)
    probe, _, pb = cpu_gcups(pkg, wl, checker, 2 * unit_n, cores, wl.seed)
    cells_per_region = pb.n_cells / pb.n_regions
    total_steps = max(1, args.steps + args.warmup)
    per_step_s = min(20.0, 150.0 / total_steps)                   # whole run within a few minutes
    n = int(max(1, min(regions, per_step_s * probe * 1e9 / cells_per_region)))
    for w in range(args.warmup):
        cpu_gcups(pkg, wl, checker, n, cores, wl.seed + w)
    cells, secs, pairs = 0, 0.0, 0
    for s in range(args.steps):
        g, dt, b = cpu_gcups(pkg, wl, checker, n, cores, wl.seed + 1000 + s)
        cells += b.n_cells
        pairs = b.n_pairs
        secs += dt
    val = cells / secs / 1e9
    sample = f"{n} {wl.key.upper()} {wl.unit} per step ({pairs} pairs), {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * secs / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name(regions), "bounded_sample": sample},
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---- e2e through phmm_submit / phmm_wait ----------------------------------------------------------
def e2e_run(pkg, eng, batches, steps, warmup, barrier=None, clk=None):
    """`steps` batches through submit/wait with DEPTH in flight; returns (seconds, h2d bytes, d2h bytes, devices used)."""
    from collections import deque
    cap = max(b.n_pairs for b in batches)                         # batches of a ragged stream differ in size
    results = [pkg.Result(cap, want_raw=False) for _ in range(2)]
    nb = len(batches)
    biggest = max(batches, key=lambda b: b.input_bytes + 8 * b.n_pairs)
    for w in range(DEPTH):                                        # every slot of the ring grows its (grow-only) buffers to the
        eng.compute(biggest, want_raw=False)                      # largest batch of the stream: no re-allocation while timed
    for w in range(warmup):
        eng.compute(batches[w % nb], want_raw=False)
    if barrier:
        barrier()
    if clk:
        clk.resume()
    t0 = time.perf_counter()
    inflight, submitted, done = deque(), 0, 0
    h2d = d2h = 0
    ndev = 0
    while done < steps:                                           # every step: H2D of its inputs, D2H of its results
        while len(inflight) < DEPTH and submitted < steps:
            inflight.append(eng.submit(batches[submitted % nb]))
            submitted += 1
        r = eng.wait(inflight.popleft(), result=results[done % 2])
        h2d += r.stats["h2d_bytes"]
        d2h += r.stats["d2h_bytes"]
        ndev = max(ndev, r.stats["n_devices_used"])
        done += 1
    if barrier:
        barrier()
    dt = time.perf_counter() - t0
    if clk:
        clk.pause()
    return dt, h2d, d2h, ndev


def pin_all(batches):
    try:
        for b in batches:
            b.pin()
        return True
    except Exception:
        for b in batches:
            b.unpin()
        return False


# ---- in-process multi-GPU: ONE engine, n_devices = N (rank 0 only) ---------------------------------
def in_process_section(pkg, n_dev, args, s5_batches):
    cores = os.cpu_count() or 1
    ht = max(2, min(4, cores // max(1, n_dev)))
    wl = WORKLOADS["s3"]
    out = {"devices": n_dev, "host_threads_per_device": ht, "host_cores": cores, "batches_in_flight": DEPTH,
           "what": "ONE host process, ONE phmm_engine with n_devices = N (worker + packer thread per device, regions of every "
                   "batch sharded over the devices by cell count, no collective); e2e through phmm_submit / phmm_wait with host buffers"}
    steps = max(8, min(args.steps, 40))

    def measure(eng, batches, label):
        rec = {}
        cells = sum(batches[i % len(batches)].n_cells for i in range(steps))
        s, h2d, d2h, ndev = e2e_run(pkg, eng, batches, steps, 3)
        rec["e2e_pageable"] = round(cells / s / 1e9, 1)
        pinned = pin_all(batches)
        if pinned:
            s, h2d, d2h, ndev = e2e_run(pkg, eng, batches, steps, 2)
            rec["e2e_pinned"] = round(cells / s / 1e9, 1)
            for b in batches:
                b.unpin()
        rec.update(devices_used=int(ndev), h2d_bytes_per_step=int(h2d / steps), d2h_bytes_per_step=int(d2h / steps),
                   steps=steps, cells_per_step=int(cells / steps), workload=label)
        return rec

    # the same engine configuration on ONE device: the denominator of the efficiency
    s3_one = [wl.make(pkg, wl.regions, 5003 + i) for i in range(4)]
    s5_one = s5_batches["one"]
    with pkg.PairHMMEngine(devices=[0], pipeline_depth=DEPTH, host_threads=ht) as e1:
        one_s3 = measure(e1, s3_one, f"S3, {wl.regions} regions per step")
        one_s5 = measure(e1, s5_one, f"S5, {s5_one[0].n_regions} windows per step")
        # whole-batch results of the N-device batches on one device: the bit-for-bit reference below
        s3_big = [wl.make(pkg, wl.regions * n_dev, 6003 + i) for i in range(3)]
        s5_big = s5_batches["big"]
        want_s3 = e1.compute(s3_big[0], want_raw=True)
        want_s5 = e1.compute(s5_big[0], want_raw=True)
    with pkg.PairHMMEngine(devices=list(range(n_dev)), pipeline_depth=DEPTH, host_threads=ht) as eN:
        got_s3 = eN.compute(s3_big[0], want_raw=True)
        got_s5 = eN.compute(s5_big[0], want_raw=True)
        par = ParityCollector()
        bitwise = True
        for b, got, want in ((s3_big[0], got_s3, want_s3), (s5_big[0], got_s5, want_s5)):
            bitwise &= bool(np.array_equal(got.log10, want.log10, equal_nan=True) and np.array_equal(got.rescued, want.rescued))
            cut = pkg.shard_bounds(b.region_cells(), n_dev)
            for d in range(n_dev):                                # 3 whole regions of every device's share
                if cut[d + 1] > cut[d]:
                    rng = np.random.default_rng(77 + d)
                    gs = sorted(set(int(g) for g in rng.integers(cut[d], cut[d + 1], 3)))
                    par.add(b, got.log10, got.rescued, gs)
        n_s3 = measure(eN, s3_big, f"S3, {wl.regions * n_dev} regions per step")
        n_s5 = measure(eN, s5_big, f"S5, {s5_big[0].n_regions} windows per step")
    pr = par.check()
    out["parity"] = pr
    out["parity_max_abs"] = None if pr is None else max(pr["max_abs_fp32"], pr["max_abs_fp64"])
    out["bitwise_equal_to_one_device_engine"] = bitwise
    for key, one, big in (("s3", one_s3, n_s3), ("s5", one_s5, n_s5)):
        rec = {"one_device": one, "n_devices": big}
        for k in ("e2e_pageable", "e2e_pinned"):
            if k in one and k in big:
                rec[f"efficiency_vs_n1_{k[4:]}"] = round(big[k] / (n_dev * one[k]), 3)
        out[key] = rec
    out["value"] = n_s3.get("e2e_pinned", n_s3["e2e_pageable"])
    out["e2e"] = n_s3["e2e_pageable"]
    out["efficiency_vs_n1"] = out["s3"].get("efficiency_vs_n1_pageable")
    out["devices_used"] = n_s3["devices_used"]
    out["ok"] = bool(bitwise and (pr is None or pr["ok"]) and n_s3["devices_used"] == n_dev)
    return out


# ---- chrM-style end-to-end wall time (BASELINE.json's second metric) -------------------------------
def e2e_chrm_section(pkg, runs=3):
    import re
    exe = lambda name: os.path.join(ROOT, "oracle", "_ref", name)
    if not (os.path.exists(exe("hc_e2e_ref")) and os.path.exists(exe("hc_e2e_b200_batched"))):
        return {"unavailable": "oracle/_ref/hc_e2e_* not built (needs /root/reference at build time)"}
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "chrm_like")
        pkg.synth.chrm_like(prefix)

        def run(name, extra=()):
            best = None
            for _ in range(runs):
                out = os.path.join(d, name + ".vcf")
                t0 = time.perf_counter()
                r = subprocess.run([exe(name), "-I", prefix + ".sam", "-R", prefix + ".fa", "-O", out, *extra],
                                   capture_output=True, text=True)
                wall = time.perf_counter() - t0
                if r.returncode != 0:
                    raise RuntimeError(r.stderr[-500:])
                m = re.search(r"init_s=([0-9.]+) do_work_s=([0-9.]+)", r.stderr)
                rec = dict(wall_s=round(wall, 3), init_s=float(m.group(1)), do_work_s=float(m.group(2)))
                if best is None or rec["wall_s"] < best["wall_s"]:
                    best = rec
            best["vcf"] = open(out).read()
            return best

        ref = run("hc_e2e_ref")
        threads = min(16, os.cpu_count() or 1)
        b200 = run("hc_e2e_b200_batched", ("-T", str(threads)))
        win = run("hc_e2e_b200")
    return {"contig": "synthetic chrM-like, 16 569 bp, <= 1 read per start (synth.chrm_like, seed 1001)", "runs_each": runs,
            "ref_wall_s": ref["wall_s"], "ref_do_work_s": ref["do_work_s"],
            "b200_wall_s": b200["wall_s"], "init_s": b200["init_s"], "do_work_s": b200["do_work_s"],
            "b200_driver": f"hc::B200RegionBatcher (cross-window batches), assembly on {threads} host threads",
            "b200_per_window_wall_s": win["wall_s"], "b200_per_window_do_work_s": win["do_work_s"],
            "vcf_identical": bool(b200["vcf"] == ref["vcf"] and win["vcf"] == ref["vcf"]),
            "wall_is": "process wall clock incl. CUDA context creation and library load; best of runs"}


# ---- our arm -----------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s3", choices=sorted(WORKLOADS))
    ap.add_argument("--regions", type=int, default=0, help="regions (windows) per GPU per step; 0 = the workload's default")
    ap.add_argument("--nbatch", type=int, default=12, help="distinct device-resident batches rotated through")
    ap.add_argument("--exact", action="store_true", help="exact_fp32 engine (raw FP32 sums bit-identical to the reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chrm", action="store_true")
    ap.add_argument("--no-sw", action="store_true")
    ap.add_argument("--no-in-process", action="store_true")
    ap.add_argument("--_helper", nargs="+", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args._helper:
        if args._helper[0] == "cpu_baseline":
            helper_cpu_baseline(args._helper[1])
        elif args._helper[0] == "parity":
            helper_parity(args._helper[1])
        elif args._helper[0] == "sw_reference":
            helper_sw_reference(args._helper[1])
        return
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3                                            # timing rule: W >= 3
    wl = WORKLOADS[args.workload]
    regions = args.regions or wl.regions

    from __graft_entry__ import load_package
    pkg = load_package()
    # Everything that forks worker processes (the S5 generator) happens BEFORE CUDA / NCCL exist in this process.
    batches = [wl.make(pkg, regions, wl.seed + 1000 * rank + i) for i in range(args.nbatch)]
    do_in_process = world > 1 and rank == 0 and not args.no_in_process
    s5_pre = None
    if do_in_process:
        s5w = WORKLOADS["s5"].regions
        s5_pre = {"one": [pkg.synth.s5_batch(s5w, seed=7005 + i) for i in range(3)],
                  "big": [pkg.synth.s5_batch(s5w * world, seed=8005 + i) for i in range(3)]}

    import torch
    torch.cuda.set_device(local)
    host_group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the one JSON line only: NCCL's banner (printed by the C library on communicator
        # creation when NCCL_DEBUG is set) goes to stderr -- file descriptor 1 points at stderr while the
        # communicator is created (eagerly, device_id given, and again under the first collective)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
            host_group = dist.new_group(backend="gloo")           # host-side waits that keep every GPU free
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    props = torch.cuda.get_device_properties(local)
    sms = props.multi_processor_count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    eng = pkg.PairHMMEngine(devices=[local], pipeline_depth=DEPTH, host_threads=4, exact_fp32=args.exact)
    # distinct batches per rank (weak scaling: every GPU gets its own `regions` regions per step)
    cells_per_step = float(np.mean([b.n_cells for b in batches]))
    eng.compute(batches[-1], want_raw=False)                       # as inside a stream: the planner's choices that follow the
    staged = [eng.stage(b) for b in batches]                       # previous batch (FP64-first for rescue-dense input) are made
    in_bytes = int(np.mean([b.input_bytes for b in batches]))
    out_bytes = int(np.mean([4 * b.n_pairs for b in batches]))
    resident_mb = sum(b.input_bytes + 20 * b.n_pairs for b in batches) / 1e6

    # ---- value: inputs resident in HBM, kernels only ----
    # (a) the dominant kernel alone: passes one after the other, CUDA events around the FP32 forward launch on
    #     its own stream (in-library) -> roofline.achieved; (b) the whole job: the K timed steps issued back
    #     to back, step i on batch i % nbatch, every batch on its own stream, so one step's FP64 redo and
    #     kernel tail overlap the next step's FP32 kernel (as behind phmm_submit with tickets in flight) ->
    #     value.  Device time from a start event every stream waits on to an end event that waits on all.
    for w in range(args.warmup):
        eng.run_staged(staged[w % args.nbatch], 1)
    k32_ms, seq_ms, k32_cells, single_launch = 0.0, 0.0, 0.0, True
    n_probe = min(args.steps, args.nbatch)
    for s in range(n_probe):
        ms, ms32, _ = eng.run_staged_ex(staged[s % args.nbatch], 1)
        seq_ms += ms
        single_launch &= ms32 > 0
        k32_ms += ms32 if ms32 > 0 else ms
        k32_cells += batches[s % args.nbatch].n_cells
    eng.run_staged_pipelined(staged, min(args.steps, args.nbatch))          # untimed: same launch pattern
    barrier()
    clk = ClockSampler(local)
    clk.__enter__()                                                # samples through BOTH timed regions (kernels, e2e)
    t0 = time.perf_counter()
    dev_ms, launches = eng.run_staged_pipelined(staged, args.steps)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clk.pause()
    dev_ms_max = max_over_ranks(dev_ms, world)
    my_cells = sum(batches[i % args.nbatch].n_cells for i in range(args.steps))
    total_cells = sum_over_ranks(my_cells, world)
    value = total_cells / (dev_ms_max * 1e-3) / 1e9
    per_gpu = k32_cells / (k32_ms * 1e-3) / 1e9                   # dominant kernel(s): cells of the launches / their duration
    per_gpu_step = k32_cells / (seq_ms * 1e-3) / 1e9              # whole steps run alone (FP32 + FP64 redo)

    # ---- parity ON THE TIMED PATH: every staged batch the timed region ran, >= 2 whole regions of each ----
    par = ParityCollector()
    n_rescued = 0
    checksum = 0.0
    per_batch = 2 if world == 1 else 1
    for i in range(min(args.steps, args.nbatch)):
        res = eng.fetch_staged(staged[i], batches[i].n_pairs, want_raw=True)
        n_rescued = int(res.stats["n_rescued"])
        checksum += float(np.sum(res.log10[np.isfinite(res.log10)]))
        par.add(batches[i], res.log10, res.rescued, pick_regions(batches[i], per_batch, 4242 + 17 * i + rank))
    parity = par.check()

    # ---- e2e: host buffers through phmm_submit / phmm_wait (H2D + kernels + D2H + log10) ----
    e2e_page_s, _, _, _ = e2e_run(pkg, eng, batches, args.steps, args.warmup, barrier, clk)
    pinned_inputs = pin_all(batches)                               # the step's inputs in page-locked host memory, uploaded
    e2e_s, h2d, d2h, _ = e2e_run(pkg, eng, batches, args.steps, args.warmup, barrier, clk)   # in place (PHMM_BATCH_PINNED_INPUTS)
    clk.__exit__()
    clocks = clk.summary()
    e2e = total_cells / max_over_ranks(e2e_s, world) / 1e9
    e2e_page = total_cells / max_over_ranks(e2e_page_s, world) / 1e9

    for st in staged:
        eng.free_staged(st)
    if pinned_inputs:
        for b in batches:
            b.unpin()
    eng.close()

    # checksum gather: proves every rank produced results (host side, after the timed regions)
    all_sums = gather_results(np.array([checksum]), rank, world, rank, world)
    all_ok = sum_over_ranks(0.0 if (parity is None or parity["ok"]) else 1.0, world) == 0.0
    worst32 = max_over_ranks(parity["max_abs_fp32"] if parity else 0.0, world)
    worst64 = max_over_ranks(parity["max_abs_fp64"] if parity else 0.0, world)
    n_checked = int(sum_over_ranks(parity["regions_checked"] if parity else 0, world))

    # ---- in-process multi-GPU scheduler: rank 0 alone, the other ranks wait on the HOST (gloo) ----
    in_process = None
    if world > 1 and not args.no_in_process:
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                in_process = in_process_section(pkg, world, args, s5_pre)
            except Exception as ex:                                # an infrastructure problem is reported, not counted as a
                in_process = {"ok": None, "error": f"{type(ex).__name__}: {ex}"}    # parity failure (only ok == False fails the run)
        dist.barrier(group=host_group)

    failed = not all_ok
    if rank == 0:
        # DRAM traffic of the dominant kernel, per launch, from the committed `ncu --set full` capture
        # (profiles/traffic.json), scaled to this launch's regions
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if args.workload == "s3":
                traffic = int((tj["dram_read_bytes"] + tj["dram_write_bytes"]) * regions / tj["regions"])
        except Exception:
            pass
        peak = sms * SM_LANES * (sm_max_mhz * 1e6) / INSTR_PER_CELL / 1e9
        mhz = clocks.get("sm_mhz") or sm_max_mhz
        peak_clk = sms * SM_LANES * (mhz * 1e6) / INSTR_PER_CELL / 1e9
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms_max / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name(regions), "regions_per_step_per_gpu": regions,
                       "pairs_per_step_per_gpu": int(np.mean([b.n_pairs for b in batches])), "cells_per_step_per_gpu": int(cells_per_step),
                       "precision_policy": ("FP32 (flush-to-zero, unfused, reference operation order: exact_fp32)" if args.exact else
                                            "FP32 (flush-to-zero, FMA)") + " with FP64 redo of pairs whose raw FP32 sum < 1e-28",
                       "rescued_pairs_last_step": n_rescued,
                       "timed_region": "K steps issued back to back, step i on batch i % nbatch, each batch on its own "
                                       "stream (consecutive steps overlap); CUDA events, max over ranks",
                       "l2_policy": f"steps rotate over {args.nbatch} distinct device-resident batches "
                                    f"({resident_mb:.0f} MB of inputs+outputs > 126 MB L2)",
                       "parallelism": f"regions sharded over {world} GPU(s), no collective on the data path"},
            "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d / args.steps),
                    "d2h_bytes_per_step": int(d2h / args.steps),
                    "path": f"phmm_submit/phmm_wait with host buffers ({'page-locked, uploaded in place' if pinned_inputs else 'staged by the engine'}), "
                            f"{DEPTH} batches in flight, 4 host threads"},
            "e2e_pageable": {"value": round(e2e_page, 1), "unit": UNIT,
                             "path": "the same with pageable numpy arrays: the engine copies them into its pinned staging first"},
            "gpu_launches": launches,
            "parity": {"max_abs_fp32": worst32, "max_abs_fp64": worst64, "regions_checked": n_checked,
                       "ok": bool(all_ok), "rank0": parity,
                       "what": "every device-resident batch the timed region ran, fetched after it; whole regions vs the CPU oracle"},
            "roofline": {"bound": "fp32_cuda_core", "achieved": round(per_gpu, 1), "peak": round(peak, 1), "unit": UNIT,
                         "frac": round(per_gpu / peak, 4), "traffic": traffic,
                         "algorithmic_bytes": in_bytes + out_bytes,
                         "peak_def": f"{sms} SMs x {SM_LANES} FP32 lanes x {sm_max_mhz:.0f} MHz / {INSTR_PER_CELL} FP32-pipe instr per cell "
                                     "(max SM clock of MEASURED_PEAKS.json; not HBM-bound: 7.6e-4 B/cell)",
                         "peak_at_clock": round(peak_clk, 1), "frac_at_clock": round(per_gpu / peak_clk, 4),
                         "executed_form": None if (args.exact or args.workload == "s3g") else {
                             "instr_per_cell": INSTR_PER_CELL_SCALED, "peak": round(peak * INSTR_PER_CELL / INSTR_PER_CELL_SCALED, 1),
                             "frac": round(per_gpu / (peak * INSTR_PER_CELL / INSTR_PER_CELL_SCALED), 4),
                             "what": "the default engine carries X / pMX and Y / pMY and folds pMM into the priors (scaled recurrence): 5 FP32-pipe instructions per cell "
                                     "instead of the 8 of the reference's expression that `peak` assumes, so `frac` can exceed 1; this is the "
                                     "fraction of the issue-rate bound of the recurrence actually executed"},
                         "kernel": ("the FP32 forward launch of the batch (forward_kernel<PolicyF32x2, ...>), timed alone with CUDA events "
                                    "on its stream, in-library (phmm_run_staged_ex)") if single_launch else
                                   "all forward launches of a step (several shapes on forked streams), CUDA events in-library",
                         "kernel_ms": round(k32_ms / n_probe, 4),
                         "step_alone": {"ms": round(seq_ms / n_probe, 4), "gcups": round(per_gpu_step, 1),
                                        "what": "one step run by itself: FP32 launch(es) + FP64 redo launch(es)"},
                         "hbm_staging_gbs": round((in_bytes + out_bytes) / (dev_ms / args.steps * 1e-3) / 1e9, 2)},
            "clocks": clocks,
            "wall_ms_kernel_region": round(wall_ms, 2),
            "checksums": None if all_sums is None else [round(float(x), 3) for x in np.atleast_1d(all_sums)],
        }
        if in_process is not None:
            line["in_process"] = in_process
            failed |= in_process.get("ok") is False
        if world == 1 and not args.no_chrm:
            try:
                line["e2e_chrm"] = e2e_chrm_section(pkg)
            except Exception as ex:
                line["e2e_chrm"] = {"unavailable": f"failed: {ex}"}
        if world == 1 and not args.no_chrm and args.workload == "s3":
            exe = os.path.join(ROOT, "tools", "cpp_surface_bench")
            try:                                                   # the reference-facing C++ surface, from std::string reads
                r = subprocess.run([exe, "1024", "64"], capture_output=True, text=True, timeout=300)
                line["e2e_cpp_surface"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"unavailable": r.stderr[-300:]}
            except Exception as ex:
                line["e2e_cpp_surface"] = {"unavailable": f"{ex}"}
        if world == 1 and not args.no_sw:
            try:
                line["sw"] = sw_section(pkg, sms, sm_max_mhz)
            except Exception as ex:
                line["sw"] = {"unavailable": f"failed: {ex}"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = run_helper("cpu_baseline", args.workload)
            except Exception as ex:                                 # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if failed:
        sys.stderr.write("bench.py: PARITY FAILURE (see the `parity` / `in_process` objects of the line)\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
