#!/bin/bash
# whole GPU suite + smoke + default bench after the entry-major plumbing (prep kernel, update io, read flag)
tag=${1:-r2t}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider --tb=short > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log; tail -2 $out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json | head -c 3000
