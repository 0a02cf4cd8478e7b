#!/bin/bash
# round 2 measurement call (1 GPU): every GPU test, smoke, the driver's bench line, the reference arm, the model clip,
# configs[2] as written (2000 frames), kernel-level timing, ncu launch list + full captures, sanitizer on the tcgen05 tests
tag=${1:-r2f}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --durations=10 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err ) 2> $out/bench_ref.time; echo "bench ref rc=$?"
timeout 500 python bench.py --workload 480p-model-clip --steps 2 --warmup 2 > $out/bench_model.json 2> $out/bench_model.err; echo "bench model rc=$?"
timeout 500 python bench.py --workload 1080p-2obj-bank-at-capacity --frames 2000 > $out/bench_1080p_2000.json 2> $out/bench_1080p_2000.err; echo "bench 1080p rc=$?"
timeout 300 python bench.py --workload 1080p-2obj-bank-at-capacity --no-cpu-baseline --no-torch-baseline > $out/bench_1080p_30.json 2> $out/bench_1080p_30.err
for cfg in "5000 1620" "100000 1620" "100000 8160"; do timeout 300 python tests/profile_kernels.py $cfg 5; done > $out/profile_kernels.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-baseline > $out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'tc_scan_pair|tc_phase_b_pair|urr_local_stream|compact_move|append_rows_warp|merge_runs' -s 6 -c 8 -f -o $out/prof \
    python tests/profile_kernels.py 100000 1620 1 > $out/ncu_full.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
    -k "read_golden or match_ties or update_golden_teacher_forced" > $out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
    -k "read_golden" > $out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $out/sanitizer_racecheck.log
cp gpurun_out/*.json $out/ 2>/dev/null
grep -E "passed|failed|FAILED|free-running|usage counts|capacity parity" $out/pytest_gpu.log | tail -20; tail -2 $out/smoke.log
cat $out/bench.json; cat $out/bench_ref.json; cat $out/bench_ref.time; cat $out/bench_model.json; cat $out/bench_1080p_2000.json; cat $out/profile_kernels.log
tail -3 $out/sanitizer_memcheck.log $out/sanitizer_racecheck.log; tail -3 $out/*.err
