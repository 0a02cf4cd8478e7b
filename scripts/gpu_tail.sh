#!/bin/bash
# Tail-first GPU check: new tail tests, then the rest of the GPU suite, smoke, one bench line.
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_tail.sh <tag>'
tag=${1:-t}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_tail.py -m gpu -q --tb=short -p no:cacheprovider > $out/pytest_tail.log 2>&1; echo "tail rc=$?" | tee -a $out/pytest_tail.log
timeout 500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_tail.py > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 400 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -30 $out/pytest_tail.log; tail -5 $out/pytest_gpu.log; tail -3 $out/smoke.log; cat $out/bench.json; tail -5 $out/bench.err
