#!/bin/bash
# final evidence of the round: ncu launch list of the default bench command + --set full of the hot kernels at capacity
tag=${1:-r3d}
out=gpurun_out/$tag
mkdir -p $out
VFN_BENCH_NO_SAMPLER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline > $out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
gzip -f $out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_scan|tc_phase_b|urr_local" -s 1 -c 3 -f -o $out/prof \
    python tests/profile_kernels.py 100000 1620 1 > $out/ncu_full.log 2>&1; echo "full rc=$?"
tail -3 $out/ncu_full.log
