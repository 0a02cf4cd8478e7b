#!/bin/bash
# 2-GPU call: the whole GPU suite (with the two-device test and the hypothesis tests), model-clip bench, 64 streams on ONE GPU
tag=${1:-r2i}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -25 $out/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 500 python bench.py --workload 480p-model-clip --steps 2 --warmup 2 > $out/bench_model.json 2> $out/bench_model.err; echo "bench model rc=$?"
cat $out/bench_model.json
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --workload 480p-64-streams --steps 1 --warmup 1 > $out/streams64_g1.json 2> $out/streams64_g1.err; echo "streams rc=$?"
cat $out/streams64_g1.json; grep -v -i warn $out/*.err | tail -5
