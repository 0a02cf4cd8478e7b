#!/bin/bash
# round 2, first GPU call: all GPU tests (incl. the whole-model drop-in, capacity-size oracle and free-running clip tests),
# smoke, the default bench line, the model-clip line.
# usage: gpurun --timeout 1700 -- 'bash scripts/gpu_r2a.sh <tag>'
tag=${1:-r2a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
nproc > $out/nproc.txt; free -g >> $out/nproc.txt; lscpu | head -25 >> $out/nproc.txt; nvidia-smi topo -m >> $out/nproc.txt 2>&1
timeout 1100 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=15 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 500 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --workload 480p-model-clip --steps 2 --warmup 1 > $out/bench_model.json 2> $out/bench_model.err; echo "bench model rc=$?"
cp gpurun_out/*.json $out/ 2>/dev/null
tail -60 $out/pytest_gpu.log; tail -3 $out/smoke.log; cat $out/bench.json; tail -5 $out/bench.err; cat $out/bench_model.json; tail -5 $out/bench_model.err
