#!/bin/bash
# ncu --set full of the hot kernels at capacity (regime C, N = 100000): usage: gpurun -- 'bash scripts/gpu_ncu.sh <tag> [regex]'
tag=${1:-ncu}
rx=${2:-tc_scan|tc_phase_b|urr_local}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 1 -c 3 -f -o $out/prof \
    python tests/profile_kernels.py 100000 1620 1 > $out/ncu_full.log 2>&1
tail -12 $out/ncu_full.log
