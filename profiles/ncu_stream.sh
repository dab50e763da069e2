CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --count-reads 0"
$CMD > gpurun_out/r01i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tf23_stream -s 3 -c 1 -o gpurun_out/r01i_tf23 -f $CMD > gpurun_out/r01i_ncu_tf23.log 2>&1
tail -c 600 gpurun_out/r01i_plain.log
