#!/bin/bash
# sharded sanity after the symmetric-memory cleanup (2 GPUs) + ncu --set full of the update's HBM-bound kernels (GPU 0)
tag=${1:-r2o}
out=gpurun_out/$tag
mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tests/multi_gpu_sharded_bench.py --frames 8 --no-4k > $out/sharded.log 2>&1; echo "sharded rc=$?"; grep '^{' $out/sharded.log | cut -c1-600
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'compact_move|append_rows_warp|evict_plan|merge_runs|plan_kernel|match_rescore' -s 12 -c 12 -f -o $out/prof_update \
    python bench.py --workload 1080p-2obj-bank-at-capacity --steps 1 --warmup 1 --frames 6 --no-cpu-baseline --no-torch-baseline > $out/ncu_update.log 2>&1; echo "ncu rc=$?"
tail -3 $out/ncu_update.log | cut -c1-300
