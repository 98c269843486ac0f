#!/bin/bash
# compute-sanitizer pass over every kernel family (tools/sanitize_target.py).  ONE tool per gpurun call
# (B200_PROFILING.md): tools/sanitize.sh memcheck | racecheck | initcheck | synccheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitize_$TOOL.txt
python tools/sanitize_target.py > gpurun_out/r02_sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02_sanitize_plain.log; exit 1; }
EXTRA=""
[ "$TOOL" = "initcheck" ] && EXTRA="--track-unused-memory no"
timeout 1500 compute-sanitizer --tool $TOOL $EXTRA --print-limit 40 --error-exitcode 0 python tools/sanitize_target.py > $OUT 2>&1
echo "exit $?" >> $OUT
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE TARGET DONE|exit " $OUT
