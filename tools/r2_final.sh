#!/bin/bash
# Round-end check of the committed tree on one B200: GPU tests, smoke, both bench arms.
mkdir -p gpurun_out; O=gpurun_out/${1:-r2fin}
python -m pytest tests -m gpu -x -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?" >> ${O}_smoke.log
python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?" >> ${O}_bench.err
python bench.py --impl reference > ${O}_bench_ref.json 2> ${O}_bench_ref.err
