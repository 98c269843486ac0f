#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/r2z
for w in s4 s5; do
  python bench.py --workload $w --steps 20 --no-chrm --no-sw > ${O}_bench_$w.json 2> ${O}_bench_$w.err
done
