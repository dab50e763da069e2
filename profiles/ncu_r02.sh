#!/usr/bin/env bash
# Round-2 ncu captures (B200_PROFILING.md recipe: the plain run must exit 0 right before each ncu run; one GPU).
#   bash profiles/ncu_r02.sh <tag>        -> gpurun_out/<tag>_{tf23,c5emit,c5sort}.ncu-rep + launch list
set -u
TAG="${1:-r02}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --count-reads 5000000 --configs="
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tf23_filter -s 3 -c 1 -o gpurun_out/${TAG}_tf23 -f $CMD > gpurun_out/${TAG}_ncu_tf23.log 2>&1
# the direct kernel (no front filter): what hit-dominated batches run
AIX_INDEX23_FILTER=off $CMD > gpurun_out/${TAG}_plain3.log 2>&1 &&
AIX_INDEX23_FILTER=off ncu --set full --clock-control none --import-source on -k regex:tf23_stream -s 3 -c 1 -o gpurun_out/${TAG}_tf23direct -f $CMD > gpurun_out/${TAG}_ncu_tf23direct.log 2>&1
# positions build at 10 % of C5 (755 M windows, 640 M keys): emit pass, one digit pass of the sort, finalize
CMD="python bench_configs.py --configs c5 --scale 0.1 --no-checks"
$CMD > gpurun_out/${TAG}_c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"positions_emit|positions_finalize" -c 2 -o gpurun_out/${TAG}_c5emit -f $CMD > gpurun_out/${TAG}_ncu_c5emit.log 2>&1
$CMD > gpurun_out/${TAG}_c5_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rs_pass|rs_hist" -s 8 -c 3 -o gpurun_out/${TAG}_c5sort -f $CMD > gpurun_out/${TAG}_ncu_c5sort.log 2>&1
ls -la gpurun_out/ | tail -12
# C1 / C4 kernels at 5 % of the BASELINE sizes (traffic per unit for bench.py's roofline.traffic)
for spec in "c1:tf13_stream_kernel:3:1" "c4:coverage_kernel:1:1"; do
  IFS=: read cfg pat skip cnt <<< "$spec"
  CMD="python bench_configs.py --configs $cfg --scale 0.05 --no-checks"
  $CMD > gpurun_out/${TAG}_${cfg}_plain.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"$pat" -s $skip -c $cnt -o gpurun_out/${TAG}_${cfg} -f $CMD > gpurun_out/${TAG}_ncu_${cfg}.log 2>&1
done
