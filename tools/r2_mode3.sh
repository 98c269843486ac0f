#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/r2m3
python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
python tools/fuzz_gpu.py 90 4242 > ${O}_fuzz.json 2> ${O}_fuzz.err; echo "fuzz rc=$?" >> ${O}_fuzz.err
python tools/bench_shapes.py > ${O}_shapes.json 2> ${O}_shapes.err
PHMM_REFERENCE_ORDER=1 python tools/bench_shapes.py > ${O}_shapes_reforder.json 2> ${O}_shapes_reforder.err
