#!/usr/bin/env bash
# ncu --set full of the C1 / C4 / C5 kernels at 5 % of the BASELINE sizes; summarised on the box (the reports are large)
TAG="${1:-r01}"
for spec in "c1:tf13_stream_kernel:3:1" "c4:coverage_kernel:1:1" "c5:positions_scan_kernel|sort_small_kernel|positions_query_kernel:0:6"; do
  IFS=: read cfg pat skip cnt <<< "$spec"
  CMD="python tests/bench_configs.py --configs $cfg --scale 0.05"
  $CMD > gpurun_out/${TAG}_${cfg}_plain.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"$pat" -s $skip -c $cnt -o /tmp/${TAG}_${cfg} -f $CMD > gpurun_out/${TAG}_${cfg}_ncu.log 2>&1
  python profiles/summarize_ncu.py kernel /tmp/${TAG}_${cfg}.ncu-rep gpurun_out/${TAG}_${cfg}_ncu.txt > /dev/null 2>&1
done
ls -la gpurun_out | tail -8
