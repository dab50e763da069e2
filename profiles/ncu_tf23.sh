#!/usr/bin/env bash
# ncu captures of the two kernels of the batch tf path (filter kernel = the headline, direct kernel = hit-dominated batches)
# after a change to either (B200_PROFILING.md recipe: the plain run exits 0 right before each ncu run; one GPU).
#   bash profiles/ncu_tf23.sh <tag>   -> gpurun_out/<tag>_tf23{,direct}_{ncu,instr}.txt, <tag>_tf23.ncu-rep, <tag>_tf23_src.csv
set -u
TAG="${1:-r02}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --count-reads 0 --configs="
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tf23_filter -s 3 -c 1 -o gpurun_out/${TAG}_tf23 -f $CMD > gpurun_out/${TAG}_ncu_tf23.log 2>&1
AIX_INDEX23_FILTER=off $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
AIX_INDEX23_FILTER=off ncu --set full --clock-control none --import-source on -k regex:tf23_stream -s 3 -c 1 -o gpurun_out/${TAG}_tf23direct -f $CMD > gpurun_out/${TAG}_ncu_tf23direct.log 2>&1
for r in tf23 tf23direct; do
  [ -f gpurun_out/${TAG}_$r.ncu-rep ] && python profiles/summarize_ncu.py kernel gpurun_out/${TAG}_$r.ncu-rep gpurun_out/${TAG}_${r}_ncu.txt > /dev/null 2>&1
done
python profiles/instr_breakdown.py gpurun_out/${TAG}_tf23.ncu-rep aindex_b200/csrc/_obj/tf_query.o _ZN3aix19tf23_filter3_kernelILi4ELi16ELb${WIN:-1}E 100000000 gpurun_out/${TAG}_tf23_instr.txt > /dev/null 2>&1
python profiles/instr_breakdown.py gpurun_out/${TAG}_tf23direct.ncu-rep aindex_b200/csrc/_obj/tf_query.o _ZN3aix18tf23_stream_kernelILi0ELb1ELi1E 100000000 gpurun_out/${TAG}_tf23direct_instr.txt > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_tf23.ncu-rep --page source --csv > gpurun_out/${TAG}_tf23_src.csv 2>/dev/null
rm -f gpurun_out/${TAG}_tf23direct.ncu-rep
du -sh gpurun_out
