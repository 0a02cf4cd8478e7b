#!/bin/bash
# Quick GPU check: parity tests (all failures listed), smoke, kernel timings, one bench line.
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_check.sh <tag>'
tag=${1:-chk}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
for n in 1620 25000 100000; do timeout 300 python tests/profile_kernels.py $n 1620 5; done > $out/profile_kernels.log 2>&1
echo "--- single-CTA phase B (VFN_PAIR=0)" >> $out/profile_kernels.log
VFN_PAIR=0 timeout 300 python tests/profile_kernels.py 100000 1620 5 >> $out/profile_kernels.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -40 $out/pytest_gpu.log; tail -3 $out/smoke.log; cat $out/profile_kernels.log; cat $out/bench.json; tail -5 $out/bench.err
