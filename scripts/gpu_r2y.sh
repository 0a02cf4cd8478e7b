#!/bin/bash
# 2 GPUs: sharded-bank tests (the two-device tests run here), sharded parity / bench at 480p, default bench at N = 2
tag=${1:-r2y}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider > $out/pytest_sharded.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_sharded.log
tail -6 $out/pytest_sharded.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tests/multi_gpu_sharded_bench.py --frames 8 --no-4k > $out/sharded.log 2>&1; echo "sharded rc=$?"; grep '^{' $out/sharded.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench n2 rc=$?"; cut -c1-700 $out/bench_n2.json
