#!/bin/bash
# round 2, second GPU call: the reworked drop-in / free-running tests, URR-local variants
tag=${1:-r2b}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -s --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
for mode in 1 2; do echo "URR mode $mode"; VFN_URR_MODE=$mode timeout 200 python tests/profile_kernels.py 5000 1620 20 | grep -E "urr|N="; done > $out/urr_modes.log 2>&1
cp gpurun_out/*.json $out/ 2>/dev/null
grep -E "free-running|usage counts vs|passed|failed|FAILED|Error|assert" $out/pytest_gpu.log | head -40; tail -30 $out/pytest_gpu.log; cat $out/urr_modes.log
