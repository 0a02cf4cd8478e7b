#!/bin/bash
tag=${1:-r2c}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python tests/debug_readout_precision.py 12 > $out/readout_precision.log 2>&1; echo "rc=$?"
grep -v Warn $out/readout_precision.log | tail -60
