#!/bin/bash
# Short GPU check: parity tests + one bench line (no kernel sweeps, no ncu).
# usage: gpurun --timeout 700 -- 'bash scripts/gpu_quick.sh <tag>'
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
timeout 500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 400 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
tail -40 $out/pytest_gpu.log; tail -3 $out/smoke.log; cat $out/bench.json; tail -5 $out/bench.err
