#!/usr/bin/env bash
# A/B of the filter kernels of the batch tf path: AIX_FILTER_KERNEL = 1 (filter word prefetched a tile ahead), 2 (two queries
# per lane), 3 (word loaded a tile ahead, loop unrolled by two), register budgets and tiles per warp.
# Usage (GPU box): bash profiles/filter_sweep2.sh <tag> [settings ...]  -> gpurun_out/<tag>_filter_sweep2.txt
TAG="${1:-sweep}"
OUT=gpurun_out/${TAG}_filter_sweep2.txt
: > $OUT
run() {
  echo "== $*" >> $OUT
  env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --count-reads 0 --configs= 2>gpurun_out/${TAG}_filter_sweep2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  Q1 %.2f Gq/s (%.3f ms)   Q2 %.2f Gq/s   direct %.2f' % (d['value']/1e9, d['ms_per_step'], d['extra']['tf23_q2_half_hits']['value']/1e9, d['roofline']['direct_kernel']['achieved_gq_s']))" >> $OUT
}
# DRAM bytes and L2 hit rate of one launch of the filter kernel under the same settings (single-pass ncu metrics)
dram() {
  env "$@" ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:tf23_filter -s 3 -c 1 --csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --count-reads 0 --configs= 2>/dev/null | grep -E "dram__|lts__|gpu__time" | awk -F'","' '{printf "  %s %s %s\n", $(NF-2), $(NF-1), $NF}' | tr -d '"' >> $OUT
}
if [ "${DRAM:-0}" = "1" ]; then shift; for cfg in "$@"; do run $cfg; dram $cfg; done; cat $OUT; exit 0; fi
if [ $# -gt 1 ]; then shift; for cfg in "$@"; do run $cfg; done; cat $OUT; exit 0; fi
run AIX_FILTER_KERNEL=1
run AIX_FILTER_KERNEL=3
run AIX_FILTER_KERNEL=3 AIX_FILTER_MINBLOCKS=4
run AIX_FILTER_KERNEL=3 AIX_FILTER_TILES=32
run AIX_FILTER_KERNEL=3 AIX_FILTER_PERSIST=1
run AIX_FILTER_KERNEL=3 AIX_FILTER_PERSIST=1 AIX_BLOOM_BITS=10
cat $OUT
