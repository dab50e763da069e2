#!/usr/bin/env bash
# ncu --set full of the coverage (C4) and positions (C5) kernels at 5 % of the BASELINE sizes
TAG="${1:-r01}"
CMD="python profiles/bench_configs.py --configs c4,c5 --scale 0.05"
$CMD > gpurun_out/${TAG}_cfg_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"coverage_kernel|positions_scan_kernel|positions_query_kernel|sort_small" -c 8 -o gpurun_out/${TAG}_cfg -f $CMD > gpurun_out/${TAG}_cfg_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_cfg_ncu.log
