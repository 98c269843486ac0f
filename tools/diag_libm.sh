#!/bin/bash
mkdir -p gpurun_out
{
PHMM_TRACE_INIT=1 python - <<'PY'
import sys; sys.path.insert(0,'.')
from __graft_entry__ import load_package
pkg=load_package()
import ctypes
L=pkg.lib()
f=getattr(L,'_ZN4phmm30log10_restatement_matches_libmEv'); f.restype=ctypes.c_bool
e=pkg.PairHMMEngine(devices=[0])
print("selftest after create:", f())
b=pkg.synth.s3(1)
r=e.compute(b)
print("launches", r.stats["kernel_launches"])
PY
} > gpurun_out/diag_libm3.txt 2>&1
