#!/bin/bash
# N-GPU check (gpurun --gpus N): sharded read + update over NCCL vs the single-GPU bank, then the stream-parallel bench.
# usage: gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_multi.sh <tag> 2'
tag=${1:-multi}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py > $out/multi_gpu_check.log 2>&1; echo "multi_gpu_check rc=$?" | tee -a $out/multi_gpu_check.log
timeout 300 python bench.py --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $n --steps 3 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench rc=$?"
grep -v "^W\|^\*\*\*" $out/multi_gpu_check.log | tail -8
for f in bench_n1 bench_n$n; do python - <<PY
import json
d=json.loads(open('$out/$f.json').read().strip().splitlines()[-1])
print('$f', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms_steps', d.get('ms_steps'), 'ranks', d['ms_per_rank'], d['clocks'])
PY
done
