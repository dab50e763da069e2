CMD="python profiles/count13_sweep.py 4000000 repetitive"
$CMD > gpurun_out/r01z_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:count13_kernel -s 12 -c 4 -o gpurun_out/r01z_count13rep -f $CMD > gpurun_out/r01z_ncu.log 2>&1
tail -3 gpurun_out/r01z_ncu.log
