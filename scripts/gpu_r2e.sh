#!/bin/bash
# 2-GPU call: sharded bank (after the gather fix), drop-in tests (fp32-conv gating), 64-streams workload at G=1 and G=2
tag=${1:-r2e}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tests/multi_gpu_sharded_bench.py --frames 24 > $out/sharded.log 2>&1; echo "sharded rc=$?"
grep -v Warning $out/sharded.log | tail -4
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider -s > $out/pytest_dropin.log 2>&1; echo "dropin rc=$?"
tail -30 $out/pytest_dropin.log
timeout 600 python bench.py --workload 480p-64-streams --total-streams 16 --frames 100 --steps 1 --warmup 1 > $out/streams_g1.json 2> $out/streams_g1.err; echo "streams g1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload 480p-64-streams --total-streams 16 --frames 100 --steps 1 --warmup 1 > $out/streams_g2.json 2> $out/streams_g2.err; echo "streams g2 rc=$?"
cat $out/streams_g1.json $out/streams_g2.json; tail -5 $out/streams_g1.err $out/streams_g2.err
cp gpurun_out/*.json $out/ 2>/dev/null
