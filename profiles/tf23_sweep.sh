#!/usr/bin/env bash
# A/B sweep of the 23-mer lookup variants (kernel shape, MPHF record shape, fingerprint bits).
# Usage (GPU box): bash profiles/tf23_sweep.sh <tag> [full]  -> gpurun_out/<tag>_sweep.txt
TAG="${1:-sweep}"
OUT=gpurun_out/${TAG}_sweep.txt
: > $OUT
run() {
  echo "== $*" >> $OUT
  env "$@" python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --count-reads 0 --configs= 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  Q1 %.2f Gq/s (%.3f ms)   Q2 %.2f Gq/s' % (d['value']/1e9, d['ms_per_step'], d['extra']['tf23_q2_half_hits']['value']/1e9))" >> $OUT
}
if [ "${2:-}" = "full" ]; then
run AIX_TF23_KERNEL=0 AIX_MPHF_WIDE=1 AIX_FP_TIER_BITS=8
run AIX_TF23_KERNEL=0 AIX_MPHF_WIDE=0 AIX_FP_TIER_BITS=4
run AIX_TF23_KERNEL=0 AIX_MPHF_WIDE=0 AIX_FP_TIER_BITS=0
run AIX_TF23_KERNEL=1 AIX_TF23_MINBLOCKS=6 AIX_MPHF_WIDE=1 AIX_FP_TIER_BITS=8
run AIX_TF23_KERNEL=1 AIX_TF23_MINBLOCKS=6 AIX_MPHF_WIDE=0 AIX_FP_TIER_BITS=0
fi
run AIX_TF23_KERNEL=0 AIX_MPHF_WIDE=0 AIX_FP_TIER_BITS=8
run AIX_TF23_KERNEL=1 AIX_TF23_MINBLOCKS=1 AIX_MPHF_WIDE=0 AIX_FP_TIER_BITS=8
run AIX_TF23_KERNEL=1 AIX_TF23_MINBLOCKS=1 AIX_MPHF_WIDE=0 AIX_FP_TIER_BITS=4
# round 2: fingerprints inside the MPHF records (3 scattered requests per query instead of 4)
run AIX_TF23_KERNEL=0 AIX_INDEX23_LAYOUT=fused
run AIX_TF23_KERNEL=1 AIX_TF23_MINBLOCKS=1 AIX_INDEX23_LAYOUT=fused
run AIX_TF23_KERNEL=1 AIX_TF23_MINBLOCKS=5 AIX_INDEX23_LAYOUT=fused
run AIX_TF23_KERNEL=1 AIX_TF23_MINBLOCKS=6 AIX_INDEX23_LAYOUT=fused
cat $OUT
