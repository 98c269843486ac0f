for mode in LAZY EAGER; do
CUDA_MODULE_LOADING=$mode PHMM_TRACE_INIT=1 python -c "
import time; t=time.perf_counter()
from __graft_entry__ import load_package
pkg=load_package(); pkg.lib(); t1=time.perf_counter()
e=pkg.PairHMMEngine(devices=[0], pipeline_depth=4, host_threads=4); t2=time.perf_counter()
b=pkg.synth.s3(1); t3=time.perf_counter(); e.compute(b); t4=time.perf_counter(); e.compute(b); t5=time.perf_counter()
print(f'CUDA_MODULE_LOADING=$mode: load lib {t1-t:.3f}s create {t2-t1:.3f}s first compute {t4-t3:.3f}s second {t5-t4:.3f}s')
" 2>&1 | grep -E "CUDA_MODULE|primary context"
done
