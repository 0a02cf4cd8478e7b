#!/bin/bash
# 2-GPU call: sharded bank over peer memory (parity + 4K measurement); drop-in tests with the exact arm on GPU 0
tag=${1:-r2d}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tests/multi_gpu_sharded_bench.py --frames 24 > $out/sharded.log 2>&1; echo "sharded rc=$?"
tail -25 $out/sharded.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider -s > $out/pytest_dropin.log 2>&1; echo "dropin rc=$?"
tail -30 $out/pytest_dropin.log
cp gpurun_out/*.json $out/ 2>/dev/null
