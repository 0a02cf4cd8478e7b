#!/bin/bash
# KeyValue head after the row-shared A operand and the staged epilogue: tests, call times, ncu launch list + full capture
tag=${1:-r3a}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_keyvalue.py -m gpu -q -p no:cacheprovider --tb=short > $out/pytest_kv.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_kv.log
timeout 300 python tests/debug_cnn_times.py > $out/cnn_times.json 2> $out/cnn_times.err; echo "probe rc=$?"; grep vfn_keyvalue $out/cnn_times.json
KV_REPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/kv_launches.csv python tests/profile_keyvalue.py > $out/kv_launches.log 2>&1; echo "launch list rc=$?"
KV_REPS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:kv_gemm -c 4 -f -o $out/kv_gemm python tests/profile_keyvalue.py > $out/kv_full.log 2>&1; echo "full rc=$?"
