#!/bin/bash
# 2-GPU pass: multi-device tests, the torchrun bench with the in-process section
TAG=${1:-r2c}; N=${2:-2}
mkdir -p gpurun_out; O=gpurun_out/$TAG
nvidia-smi -L > ${O}_gpus.txt
python -m pytest tests -m gpu -x -q -k "multi_device or sharded or full_size or pinned_inputs" > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > ${O}_bench_n$N.json 2> ${O}_bench_n$N.err; echo "bench rc=$?" >> ${O}_bench_n$N.err
PHMM_TRACE_INIT=1 python -c "
import time; t=time.perf_counter()
from __graft_entry__ import load_package
pkg=load_package(); pkg.lib(); t1=time.perf_counter()
e=pkg.PairHMMEngine(devices=[0]); t2=time.perf_counter()
b=pkg.synth.s3(1); t3=time.perf_counter(); e.compute(b); t4=time.perf_counter(); e.compute(b); t5=time.perf_counter()
print(f'load lib {t1-t:.3f}s create {t2-t1:.3f}s first compute {t4-t3:.3f}s second {t5-t4:.3f}s')
" > ${O}_init.txt 2>&1
ls -la gpurun_out
