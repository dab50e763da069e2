#!/usr/bin/env bash
# Round-2 ncu captures (B200_PROFILING.md recipe: the plain run must exit 0 right before each ncu run; one GPU).
#   bash profiles/ncu_r02.sh <tag>        -> gpurun_out/<tag>_*_ncu.txt summaries (+ the report of the headline kernel) + launch list
# gpurun brings back at most 64 MiB: the reports are summarised on the box and all but the headline kernel's are deleted.
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

# summaries on the box (the reports together exceed what gpurun_out may carry back)
for r in tf23 tf23direct c5emit c5sort c1 c4; do
  [ -f gpurun_out/${TAG}_$r.ncu-rep ] && python profiles/summarize_ncu.py kernel gpurun_out/${TAG}_$r.ncu-rep gpurun_out/${TAG}_${r}_ncu.txt > /dev/null 2>&1
done
python profiles/summarize_ncu.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.txt > /dev/null 2>&1
python profiles/instr_breakdown.py gpurun_out/${TAG}_tf23.ncu-rep aindex_b200/csrc/_obj/tf_query.o _ZN3aix19tf23_filter3_kernelILi4ELi16ELb1E 100000000 gpurun_out/${TAG}_tf23_instr.txt > /dev/null 2>&1
python profiles/instr_breakdown.py gpurun_out/${TAG}_tf23direct.ncu-rep aindex_b200/csrc/_obj/tf_query.o _ZN3aix18tf23_stream_kernelILi0ELb1ELi1E 100000000 gpurun_out/${TAG}_tf23direct_instr.txt > /dev/null 2>&1
python profiles/instr_breakdown.py gpurun_out/${TAG}_c5sort.ncu-rep aindex_b200/csrc/_obj/radix_sort.o _ZN3aix14rs_pass_kernelILi256ELi7ELi4ENS_9BitsDigitE 640000000 gpurun_out/${TAG}_c5sort_instr.txt > /dev/null 2>&1
rm -f gpurun_out/${TAG}_tf23direct.ncu-rep gpurun_out/${TAG}_c5emit.ncu-rep gpurun_out/${TAG}_c1.ncu-rep gpurun_out/${TAG}_c4.ncu-rep gpurun_out/${TAG}_launches.csv
du -sh gpurun_out
