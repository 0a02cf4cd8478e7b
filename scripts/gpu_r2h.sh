#!/bin/bash
# 8-GPU call: scaling of the default workload (device-timed, e2e, copy-only leg), 64 streams over 8 GPUs, sharded bank at 8 and 4 GPUs
tag=${1:-r2h}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1; nproc >> $out/topo.txt; lscpu | grep -E "Model name|Socket|NUMA|Core" >> $out/topo.txt
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29515 "$@"; }
run 8 bench.py --gpus 8 --steps 10 --warmup 3 > $out/bench_n8.json 2> $out/bench_n8.err; echo "bench n8 rc=$?"
run 8 bench.py --gpus 8 --workload 480p-64-streams --steps 1 --warmup 1 > $out/streams_n8.json 2> $out/streams_n8.err; echo "streams n8 rc=$?"
run 8 tests/multi_gpu_sharded_bench.py --frames 24 > $out/sharded_n8.log 2>&1; echo "sharded n8 rc=$?"
run 4 tests/multi_gpu_sharded_bench.py --frames 24 --no-parity > $out/sharded_n4.log 2>&1; echo "sharded n4 rc=$?"
run 4 bench.py --gpus 4 --steps 10 --warmup 3 > $out/bench_n4.json 2> $out/bench_n4.err; echo "bench n4 rc=$?"
cp gpurun_out/sharded_bench_n*.json $out/ 2>/dev/null
cat $out/bench_n8.json; cat $out/streams_n8.json; grep '^{' $out/sharded_n8.log; grep '^{' $out/sharded_n4.log; cat $out/bench_n4.json
grep -v -i warn $out/*.err | tail -10
