#!/bin/bash
# Round-2 opening GPU pass: parity tests, every shape's throughput, launch lists + full ncu of the sub-70% shapes.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python tools/bench_shapes.py > gpurun_out/r2a_shapes.json 2> gpurun_out/r2a_shapes.err
for spec in "16 s4" "64 s4" "128 s3g"; do
  set -- $spec
  python tools/prof_s3.py $1 $2 > gpurun_out/r2a_plain_$2_$1.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file gpurun_out/r2a_launches_$2_$1.csv python tools/prof_s3.py $1 $2 > /dev/null 2>&1
done
python tools/prof_s5.py 1024 > gpurun_out/r2a_plain_s5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2a_launches_s5.csv python tools/prof_s5.py 1024 > /dev/null 2>&1
# full captures: S4 (second pass: launches 8..15), MODE 0 (launches 2..3)
ncu --set full --clock-control none -s 8 -c 6 -o gpurun_out/r2a_s4_full python tools/prof_s3.py 16 s4 > gpurun_out/r2a_ncu_s4.log 2>&1
ncu --set full --clock-control none -s 2 -c 2 -o gpurun_out/r2a_s3g_full python tools/prof_s3.py 128 s3g > gpurun_out/r2a_ncu_s3g.log 2>&1
ls -la gpurun_out
