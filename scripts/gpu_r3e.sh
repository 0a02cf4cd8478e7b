#!/bin/bash
# final validation as the driver runs it (1 GPU): GPU tests, smoke, bench (reference arm + ours), model clip
tag=${1:-r3e}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log; tail -2 $out/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload 480p-model-clip --steps 2 --warmup 2 > $out/bench_model.json 2> $out/bench_model.err; echo "bench model rc=$?"
cut -c1-600 $out/bench_ref.json; cut -c1-900 $out/bench.json
