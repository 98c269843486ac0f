#!/usr/bin/env python
"""bench.py -- PairHMM forward throughput (GCUPS) on B200, the metric of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W              (our arm; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU PairHMM)

Workload (config.workload): BASELINE.json configs[2] -- synthetic 150 bp reads x 500 bp haplotypes,
256 reads x 16 haplotypes per active region (seed 1003, SURVEY.md section 8d "S3"), REGIONS regions
per GPU per step.  A step = one pass of the hot path (FP32 forward kernels + FP64 rescue kernels)
over one such batch.  GCUPS = sum over pairs of read_len x hap_len / seconds / 1e9, each pair once.

  value      inputs already resident in HBM (phmm_stage), kernels only, CUDA events on the launching
             stream (inside the C library), max over ranks, summed over ranks' cells.
  e2e        the same batches through the C ABI with HOST buffers (phmm_submit / phmm_wait, pipeline
             4 batches in flight): pack to pinned staging + H2D + kernels + D2H + host log10, wall clock.
  roofline   FP32 CUDA-core issue roofline of SURVEY.md section 8(d): SMs x 128 lanes x f_SM / 8
             FP32-pipe instructions per cell (NOT HBM: 7.6e-4 B/cell); `peak` uses the max SM clock
             of MEASURED_PEAKS.json, `peak_at_clock` the median clock sampled during the run.
  cpu_baseline  the reference's own AVX PairHMM (oracle/_ref, kind "reference") -- or the oracle port
             when that library is absent -- on this box's host cores, bounded sample, rank 0, N=1.

L2 rule: the timed steps rotate over NBATCH distinct device-resident batches whose inputs + outputs
exceed the 126 MB L2 (config.l2_policy).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "pairhmm_gcups"
UNIT = "GCUPS"
SM_LANES = 128          # FP32 lanes per SM
INSTR_PER_CELL = 8      # FP32-pipe instructions per cell with FMA
READ_LEN, HAP_LEN, READS, HAPS = 150, 500, 256, 16


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
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "power_w_max": max(self.power) if self.power else None,
                "samples": len(self.samples)}


# ---- CPU arm -----------------------------------------------------------------------------------
def cpu_checker():
    """(checker, kind): the compiled reference if oracle/_ref travelled here, else the oracle port."""
    from _oracle import load_oracle, load_ref
    ref = None
    try:
        ref = load_ref()
    except OSError:
        ref = None
    return (ref, "reference") if ref is not None else (load_oracle(), "port")


def cpu_gcups(pkg, checker, n_regions, threads, seed=1003):
    b = pkg.synth.s3(n_regions, seed=seed)
    t0 = time.perf_counter()
    checker.batch(b, threads=threads)
    dt = time.perf_counter() - t0
    return b.n_cells / dt / 1e9, dt


def cpu_baseline(pkg, budget_s=12.0):
    checker, kind = cpu_checker()
    cores = os.cpu_count() or 1
    g1, dt1 = cpu_gcups(pkg, checker, 1, 1)                       # as shipped: OpenMP compiled out
    gN_probe, _ = cpu_gcups(pkg, checker, 2, cores)
    n = int(max(2, min(1024, budget_s * gN_probe * 1e9 / (READS * HAPS * READ_LEN * HAP_LEN))))
    gN, dtN = cpu_gcups(pkg, checker, n, cores)
    return {"value": round(gN, 3), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} S3 regions ({n * READS * HAPS} pairs, {n * READS * HAPS * READ_LEN * HAP_LEN:.3e} cells) "
                      f"in {dtN:.1f} s with {cores} threads (omp dynamic over reads, intel_pairhmm.hpp:128-130)",
            "as_shipped_1_thread": round(g1, 3)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from __graft_entry__ import load_package
    pkg = load_package()
    checker, kind = cpu_checker()
    cores = os.cpu_count() or 1
    probe, _ = cpu_gcups(pkg, checker, 2, cores)
    total_steps = max(1, args.steps + args.warmup)
    per_step_s = min(20.0, 150.0 / total_steps)                   # whole run within a few minutes
    n = int(max(1, min(args.regions, per_step_s * probe * 1e9 / (READS * HAPS * READ_LEN * HAP_LEN))))
    for w in range(args.warmup):
        cpu_gcups(pkg, checker, n, cores, seed=1003 + w)
    cells, secs = 0, 0.0
    for s in range(args.steps):
        g, dt = cpu_gcups(pkg, checker, n, cores, seed=2003 + s)
        cells += n * READS * HAPS * READ_LEN * HAP_LEN
        secs += dt
    val = cells / secs / 1e9
    sample = f"{n} S3 regions per step ({n * READS * HAPS} pairs), {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * secs / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.regions), "bounded_sample": sample},
        "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_name(regions):
    return (f"S3 (BASELINE configs[2]): synthetic PairHMM batch, {READ_LEN}bp reads x {HAP_LEN}bp haplotypes, "
            f"{READS} reads x {HAPS} haplotypes per region, {regions} regions per GPU per step, seed 1003+")


# ---- our arm -----------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regions", type=int, default=128, help="S3 regions per GPU per step")
    ap.add_argument("--nbatch", type=int, default=12, help="distinct device-resident batches rotated through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3                                            # timing rule: W >= 3

    import torch
    torch.cuda.set_device(local)
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
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from __graft_entry__ import load_package
    pkg = load_package()
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

    DEPTH = 4                                                      # batches in flight on the e2e path
    eng = pkg.PairHMMEngine(devices=[local], pipeline_depth=DEPTH, host_threads=4)
    # distinct batches per rank (weak scaling: every GPU gets its own `regions` regions per step)
    batches = [pkg.synth.s3(args.regions, seed=1003 + 1000 * rank + i) for i in range(args.nbatch)]
    cells_per_step = batches[0].n_cells
    staged = [eng.stage(b) for b in batches]
    in_bytes = batches[0].input_bytes
    out_bytes = 4 * batches[0].n_pairs
    resident_mb = args.nbatch * (in_bytes + out_bytes + 16 * batches[0].n_pairs) / 1e6

    # ---- value: inputs resident in HBM, kernels only ----
    # (a) the dominant kernel alone: passes one after the other, CUDA events around the FP32 forward launch on
    #     its own stream (in-library) -> roofline.achieved; (b) the whole job: the K timed steps issued back
    #     to back, step i on batch i % nbatch, every batch on its own stream, so one step's FP64 redo and
    #     kernel tail overlap the next step's FP32 kernel (as behind phmm_submit with tickets in flight) ->
    #     value.  Device time from a start event every stream waits on to an end event that waits on all.
    for w in range(args.warmup):
        eng.run_staged(staged[w % args.nbatch], 1)
    k32_ms, seq_ms = 0.0, 0.0
    n_probe = min(args.steps, args.nbatch)
    for s in range(n_probe):
        ms, ms32, _ = eng.run_staged_ex(staged[s % args.nbatch], 1)
        seq_ms += ms
        k32_ms += ms32 if ms32 > 0 else ms
    k32_ms /= n_probe
    seq_ms /= n_probe
    eng.run_staged_pipelined(staged, min(args.steps, args.nbatch))          # untimed: same launch pattern
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        dev_ms, launches = eng.run_staged_pipelined(staged, args.steps)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = clk.summary()
    dev_ms_max = max_over_ranks(dev_ms, world)
    total_cells = sum_over_ranks(cells_per_step * args.steps, world)
    value = total_cells / (dev_ms_max * 1e-3) / 1e9
    per_gpu = cells_per_step / (k32_ms * 1e-3) / 1e9              # dominant kernel: cells of one launch / its duration
    per_gpu_step = cells_per_step / (seq_ms * 1e-3) / 1e9         # one whole step run alone (FP32 + FP64 redo)

    # parity spot check on the last batch that ran (outside the timed region)
    res = eng.fetch_staged(staged[(args.steps - 1) % args.nbatch], batches[0].n_pairs, want_raw=True)
    n_rescued = int(res.stats["n_rescued"])
    checksum = float(np.sum(res.log10[np.isfinite(res.log10)]))

    # ---- e2e: host buffers through phmm_submit / phmm_wait (pack + H2D + kernels + D2H + log10) ----
    from collections import deque
    results = [pkg.Result(b.n_pairs, want_raw=False) for b in batches[:2]]
    pinned_inputs = True
    try:                                                          # the step's inputs live in page-locked host memory and are
        for b in batches:                                         # uploaded from there (PHMM_BATCH_PINNED_INPUTS); if the
            b.pin()                                               # registration is refused, the engine stages them itself
    except Exception:
        pinned_inputs = False
        for b in batches:
            b.unpin()
    for w in range(max(args.warmup, DEPTH)):                      # also grows every slot's buffers
        eng.compute(batches[w % args.nbatch], want_raw=False)
    barrier()
    t0 = time.perf_counter()
    inflight, submitted, done = deque(), 0, 0
    h2d = d2h = 0
    while done < args.steps:                                      # every step: H2D of its inputs, D2H of its results
        while len(inflight) < DEPTH and submitted < args.steps:
            inflight.append(eng.submit(batches[submitted % args.nbatch])); submitted += 1
        r = eng.wait(inflight.popleft(), result=results[done % 2])
        h2d += r.stats["h2d_bytes"]; d2h += r.stats["d2h_bytes"]
        done += 1
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_s_max = max_over_ranks(e2e_s, world)
    e2e = total_cells / e2e_s_max / 1e9

    for st in staged:
        eng.free_staged(st)
    if pinned_inputs:
        for b in batches:
            b.unpin()
    eng.close()

    # checksum gather: proves every rank produced results (host side, after the timed regions)
    all_sums = gather_results(np.array([checksum]), rank, world, rank, world)

    if rank == 0:
        # DRAM traffic of the dominant kernel, per launch, from the committed `ncu --set full` capture
        # (profiles/r01_ncu_bench_forward_kernel.txt: a 64-region launch), scaled to this launch's regions
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = int((tj["dram_read_bytes"] + tj["dram_write_bytes"]) * args.regions / tj["regions"])
        except Exception:
            pass
        peak = sms * SM_LANES * (sm_max_mhz * 1e6) / INSTR_PER_CELL / 1e9
        mhz = clocks.get("sm_mhz") or sm_max_mhz
        peak_clk = sms * SM_LANES * (mhz * 1e6) / INSTR_PER_CELL / 1e9
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms_max / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.regions), "regions_per_step_per_gpu": args.regions,
                       "pairs_per_step_per_gpu": batches[0].n_pairs, "cells_per_step_per_gpu": cells_per_step,
                       "precision_policy": "FP32 (flush-to-zero, FMA) with FP64 redo of pairs whose raw FP32 sum < 1e-28",
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
            "gpu_launches": launches,
            "roofline": {"bound": "fp32_cuda_core", "achieved": round(per_gpu, 1), "peak": round(peak, 1), "unit": UNIT,
                         "frac": round(per_gpu / peak, 4), "traffic": traffic,
                         "algorithmic_bytes": in_bytes + out_bytes,
                         "peak_def": f"{sms} SMs x {SM_LANES} FP32 lanes x {sm_max_mhz:.0f} MHz / {INSTR_PER_CELL} FP32-pipe instr per cell "
                                     "(max SM clock of MEASURED_PEAKS.json; not HBM-bound: 7.6e-4 B/cell)",
                         "peak_at_clock": round(peak_clk, 1), "frac_at_clock": round(per_gpu / peak_clk, 4),
                         "kernel": "forward_kernel<PolicyF32x2, K=10, G=16, MODE=2, ALIGNED>: one launch scores the whole batch; "
                                   "timed alone with CUDA events on its stream, in-library (phmm_run_staged_ex)",
                         "kernel_ms": round(k32_ms, 4),
                         "step_alone": {"ms": round(seq_ms, 4), "gcups": round(per_gpu_step, 1),
                                        "what": "one step run by itself: FP32 launch + its FP64 redo launch"},
                         "hbm_staging_gbs": round((in_bytes + out_bytes) / (dev_ms / args.steps * 1e-3) / 1e9, 2)},
            "clocks": clocks,
            "wall_ms_kernel_region": round(wall_ms, 2),
            "checksums": None if all_sums is None else [round(float(x), 3) for x in np.atleast_1d(all_sums)],
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(pkg)
            except Exception as ex:                                 # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
