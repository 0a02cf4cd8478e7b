#!/bin/bash
# ncu of the KeyValue head: launch list (every kernel of a call) + one --set full capture of the GEMM
tag=${1:-r2s}
out=gpurun_out/$tag
mkdir -p $out
KV_REPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/kv_launches.csv python tests/profile_keyvalue.py > $out/kv_launches.log 2>&1; echo "launch list rc=$?"
grep -c kv_ $out/kv_launches.csv
KV_REPS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:kv_gemm -c 4 -o $out/kv_gemm python tests/profile_keyvalue.py > $out/kv_full.log 2>&1; echo "full rc=$?"
ls -la $out
