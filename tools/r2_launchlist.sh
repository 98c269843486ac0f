#!/bin/bash
# ncu launch list of the bench command itself (gpu__time_duration.sum; cold-cache, serialised: the SHARES are what
# must agree with bench.py's live CUDA-event numbers), taken after the same command exited 0 without ncu.
mkdir -p gpurun_out; O=gpurun_out/${1:-r2ll}
ARGS="--steps 2 --warmup 1 --no-chrm --no-sw"
python bench.py $ARGS > ${O}_plain.json 2> ${O}_plain.err || { echo "plain run failed" >> ${O}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches_bench.csv \
    python bench.py $ARGS > ${O}_under_ncu.json 2> ${O}_under_ncu.err
python - "$O" <<'PY'
import csv, sys, collections, json
o = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(o + "_launches_bench.csv") if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, pi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Process ID")
per = collections.OrderedDict()
for r in rows:
    name = r[ki].split("(")[0]
    name = name[:90]
    a = per.setdefault((r[pi], name), [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = collections.Counter()
for (pid, n), (c, t) in per.items(): tot[pid] += t
out = [{"pid": pid, "kernel": n, "launches": c, "total_us": round(t / 1e3, 1), "share_of_process": round(t / tot[pid], 4)} for (pid, n), (c, t) in per.items()]
json.dump(out, open(o + "_launches_bench_summary.json", "w"), indent=1)
PY
