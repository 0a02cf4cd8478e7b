#!/bin/bash
# other bench workloads after the round-2 kernel changes (sanity): 64 streams on one GPU (short), 1080p at capacity (short)
tag=${1:-r3b}
out=gpurun_out/$tag
mkdir -p $out
timeout 500 python bench.py --workload 480p-64-streams --steps 1 --warmup 1 --no-cpu-baseline > $out/streams64.json 2> $out/streams64.err; echo "streams rc=$?"; cut -c1-900 $out/streams64.json; tail -3 $out/streams64.err
timeout 500 python bench.py --workload 1080p-2obj-bank-at-capacity --steps 2 --warmup 1 --frames 30 --no-cpu-baseline --no-torch-baseline > $out/cap1080.json 2> $out/cap1080.err; echo "1080p rc=$?"; cut -c1-900 $out/cap1080.json; tail -3 $out/cap1080.err
