CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --count-reads 0"
$CMD > gpurun_out/r01b_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tf23_fixed -s 1 -c 1 -o gpurun_out/r01b_tf23 -f $CMD > gpurun_out/r01b_ncu_tf23.log 2>&1; ls -la gpurun_out | tail -4
