#!/bin/bash
TAG=${1:-r2s}; mkdir -p gpurun_out; O=gpurun_out/$TAG
python -m pytest tests -m gpu -x -q -s -k "hunt_for" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
for i in 1; do PHMM_TRACE_INIT=1 python -c "
import time; t=time.perf_counter()
from __graft_entry__ import load_package
pkg=load_package(); pkg.lib(); t1=time.perf_counter()
e=pkg.PairHMMEngine(devices=[0], pipeline_depth=4, host_threads=4); t2=time.perf_counter()
b=pkg.synth.s3(1); t3=time.perf_counter(); e.compute(b); t4=time.perf_counter(); e.compute(b); t5=time.perf_counter()
print(f'load lib {t1-t:.3f}s create {t2-t1:.3f}s first compute {t4-t3:.3f}s second {t5-t4:.3f}s')
"; done > ${O}_init.txt 2>&1

