#!/bin/bash
tag=${1:-r2g}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider -s -k "graphed or patch_model" > $out/pytest_graphed.log 2>&1; echo "graphed rc=$?"
tail -15 $out/pytest_graphed.log
timeout 500 python bench.py --workload 480p-model-clip --steps 2 --warmup 2 > $out/bench_model.json 2> $out/bench_model.err; echo "bench model rc=$?"
cat $out/bench_model.json; grep -v Warn $out/bench_model.err | tail -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
cp gpurun_out/*.json $out/ 2>/dev/null
