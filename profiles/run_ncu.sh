#!/usr/bin/env bash
# Profiling recipe (B200_PROFILING.md): the plain run must exit 0 right before each ncu run.
# Usage (on the GPU box, via gpurun): bash profiles/run_ncu.sh <tag>
set -u
TAG="${1:-r01}"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --count-reads 5000000"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tf23_fixed -s 1 -c 1 -o gpurun_out/${TAG}_tf23 -f $CMD > gpurun_out/${TAG}_ncu_tf23.log 2>&1
$CMD > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:count13_kernel -s 4 -c 2 -o gpurun_out/${TAG}_count13 -f $CMD > gpurun_out/${TAG}_ncu_count13.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log | cut -c1-600
ls -la gpurun_out/
