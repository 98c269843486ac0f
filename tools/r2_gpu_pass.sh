#!/bin/bash
# GPU pass: parity tests, the bench line, every shape's throughput, launch lists + full-set ncu summaries
# (the .ncu-rep files are summarised ON the box by tools/ncu_summary.py and deleted: gpurun_out is capped at 64 MiB).
# usage: r2_gpu_pass.sh <tag> [steps...]   steps: pytest bench shapes launches ncu_s4 ncu_s3g ncu_s5 ncu_s3 ncu_s2
TAG=$1; shift
STEPS="${@:-pytest bench shapes launches ncu_s4 ncu_s3g}"
mkdir -p gpurun_out
O=gpurun_out/$TAG
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has pytest; then python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; fi
if has bench; then python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?" >> ${O}_bench.err; fi
if has shapes; then python tools/bench_shapes.py > ${O}_shapes.json 2> ${O}_shapes.err; fi
if has launches; then
  for spec in "64 s4" "128 s3g"; do
    set -- $spec
    python tools/prof_s3.py $1 $2 > ${O}_plain_$2_$1.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file ${O}_launches_$2_$1.csv python tools/prof_s3.py $1 $2 > /dev/null 2>&1
  done
  python tools/prof_s5.py 1024 > ${O}_plain_s5.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file ${O}_launches_s5.csv python tools/prof_s5.py 1024 > /dev/null 2>&1
fi
full() {  # name, skip, count, cmd...
  local name=$1 skip=$2 cnt=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s $skip -c $cnt -o /tmp/${TAG}_$name "$@" > ${O}_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_$name.ncu-rep ${O}_ncu_$name.txt
  rm -f /tmp/${TAG}_$name.ncu-rep
}
if has ncu_s4; then full s4 20 6 python tools/prof_s3.py 64 s4; fi
if has ncu_s3g; then full s3g 4 2 python tools/prof_s3.py 128 s3g; fi
if has ncu_s3; then full s3 4 2 python tools/prof_s3.py 128 s3; fi
if has ncu_s2; then full s2 4 2 python tools/prof_s3.py 256 s2; fi
if has ncu_s5; then full s5 48 4 python tools/prof_s5.py 1024; fi
du -sh gpurun_out; ls -la gpurun_out
if has mode0exp; then
  for g in 16 32; do PHMM_FORCE_GROUP=$g python tools/prof_s3.py 128 s3g > ${O}_mode0_forceG$g.log 2>&1; done
fi
