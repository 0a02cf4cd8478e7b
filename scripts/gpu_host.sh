#!/bin/bash
# Host-overhead probe: per-stage issue times, bench with / without the clock sampler, longer timed region.
tag=${1:-host}
out=gpurun_out/$tag
mkdir -p $out
timeout 200 python tests/debug_host_overhead.py > $out/host_overhead.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err
VFN_BENCH_NO_SAMPLER=1 timeout 300 python bench.py --no-cpu-baseline > $out/bench_nosampler.json 2> $out/bench_nosampler.err
timeout 300 python bench.py --no-cpu-baseline --steps 10 > $out/bench_k10.json 2> $out/bench_k10.err
cat $out/host_overhead.log
for f in bench bench_nosampler bench_k10; do python - <<PY
import json
d=json.loads(open('$out/$f.json').read().strip().splitlines()[-1])
print('$f', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms_steps', d.get('ms_steps'), 'prof', round(d['roofline']['ms_per_step_profiled'],1), d['clocks'])
PY
done
