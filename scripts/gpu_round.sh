#!/bin/bash
# One gpurun call: GPU parity tests, kernel-level timing, bench line, ncu launch list + full capture of the hot kernels.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
for cfg in "1620 1620" "25000 1620" "100000 1620" "100000 8160"; do timeout 300 python tests/profile_kernels.py $cfg 5; done > $out/profile_kernels.log 2>&1
timeout 120 python tests/profile_tail.py 1080 1920 3 > $out/profile_tail.log 2>&1
timeout 120 python tests/debug_frame_times.py > $out/frame_times.log 2>&1
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'tc_scan|tc_phase_b|urr_local|compact_move' -s 4 -c 4 -f -o $out/prof \
    python tests/profile_kernels.py 100000 1620 1 > $out/ncu_full.log 2>&1
tail -3 $out/pytest_gpu.log; tail -2 $out/smoke.log; cat $out/profile_kernels.log; cat $out/profile_tail.log; cat $out/bench.json; cat $out/bench_ref.json
