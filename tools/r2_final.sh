#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/${1:-r2fin}
python -m pytest tests -m gpu -x -q -s > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?" >> ${O}_smoke.log
python bench.py > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?" >> ${O}_bench.err
python bench.py --impl reference > ${O}_bench_ref.json 2> ${O}_bench_ref.err
python bench.py --steps 2000 --warmup 5 --no-chrm --no-sw --no-cpu-baseline > ${O}_bench_sustained.json 2> ${O}_bench_sustained.err
for w in s2 s4 s5 s3g; do
  python bench.py --workload $w --steps 20 --no-chrm --no-sw --no-cpu-baseline > ${O}_bench_$w.json 2> ${O}_bench_$w.err
done
python bench.py --exact --steps 20 --no-chrm --no-sw --no-cpu-baseline > ${O}_bench_exact.json 2> ${O}_bench_exact.err
