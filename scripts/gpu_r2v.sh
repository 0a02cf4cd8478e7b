#!/bin/bash
# entry-major preparation (staged norm, aliased raw rows): KeyValue / glue tests, update parity tests, model-clip bench
tag=${1:-r2v}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_keyvalue.py tests/test_gpu_parity.py tests/test_properties.py -q -m gpu --tb=short -p no:cacheprovider > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest.log
tail -8 $out/pytest.log
timeout 600 python bench.py --workload 480p-model-clip --steps 2 --warmup 2 > $out/bench_model_clip.json 2> $out/bench_model_clip.err; echo "bench rc=$?"
TAG=$tag python - <<'P'
import json
d=json.load(open('gpurun_out/'+__import__('os').environ.get('TAG','r2v')+'/bench_model_clip.json'))
print(json.dumps({k:d[k] for k in ('value','stages_ms_per_frame','graphed_convolutions','fused_glue')},indent=1))
P
