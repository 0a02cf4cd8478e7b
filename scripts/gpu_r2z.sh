#!/bin/bash
# compute-sanitizer memcheck on the round-2 additions: KeyValue head kernels, entry-major preparation, fused combine
tag=${1:-r2z}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_keyvalue.py -m gpu -q -p no:cacheprovider \
   -k "golden or shapes or layouts or no_bias or entry_major" > $out/memcheck_kv.log 2>&1; echo "memcheck kv rc=$?"
tail -6 $out/memcheck_kv.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
   -k "read_golden or update_golden_teacher_forced or match_ties" > $out/memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"
tail -6 $out/memcheck_parity.log
