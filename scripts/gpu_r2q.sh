#!/bin/bash
# KeyValue head (vfn_kv.cu) + fused glue: new GPU tests, timing probe, model-clip bench with the fused legs.
# usage: gpurun --timeout 1200 -- 'bash scripts/gpu_r2q.sh <tag>'
tag=${1:-r2q}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_keyvalue.py -q --tb=short -p no:cacheprovider > $out/pytest_kv.log 2>&1; echo "pytest kv rc=$?" | tee -a $out/pytest_kv.log
tail -25 $out/pytest_kv.log
timeout 300 python tests/debug_cnn_times.py > $out/cnn_times.json 2> $out/cnn_times.err; echo "probe rc=$?"; cat $out/cnn_times.json; tail -3 $out/cnn_times.err
timeout 600 python bench.py --workload 480p-model-clip --steps 2 --warmup 2 > $out/bench_model_clip.json 2> $out/bench_model_clip.err; echo "bench rc=$?"
cat $out/bench_model_clip.json; tail -5 $out/bench_model_clip.err
