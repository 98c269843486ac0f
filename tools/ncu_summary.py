#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a short text: key counters, stall ratios, per-opcode stall samples.
usage: ncu_summary.py <report.ncu-rep> [out.txt]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
def page(p):
    txt = subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))
rows = page("raw")
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"), file=out)
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    for k in keys:
        if k in d: print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}", file=out)
    print("  stall ratios (warps per issue-active cycle):", file=out)
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(v) >= 0.02: print(f"    {k.split('issue_stalled_')[1].split('_per_issue')[0]:22s} {float(v):6.3f}", file=out)
            except ValueError: pass
rows = page("source")
if len(rows) > 2:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(collections.Counter); tot = 0
    reasons = ["stall_math", "stall_wait", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_mio", "stall_no_inst"]
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ix["# Samples"]] == "# Samples": continue   # later kernels repeat the header
        parts = r[ix["Source"]].split()
        if not parts: continue
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0] + (".MOV" if ".MOV" in op else "")
        n = int(r[ix["# Samples"]]); tot += n
        agg[op]["samples"] += n; agg[op]["exec"] += int(r[ix["Instructions Executed"]])
        for k in reasons: agg[op][k] += int(r[ix[k]])
    print(f"  per-opcode warp-stall samples (total {tot}):", file=out)
    print(f"    {'op':10s} {'samples':>8s} {'exec(M)':>8s} " + " ".join(f"{k[6:12]:>6s}" for k in reasons), file=out)
    for op, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:14]:
        print(f"    {op:10s} {c['samples']:8d} {c['exec']/1e6:8.1f} " + " ".join(f"{c[k]:6d}" for k in reasons), file=out)
