#!/usr/bin/env bash
# A/B sweep of the front filter of the batch tf path: off, bits per key, register budget of the filter kernel.
# Usage (GPU box): bash profiles/filter_sweep.sh <tag>  -> gpurun_out/<tag>_filter_sweep.txt
TAG="${1:-sweep}"
OUT=gpurun_out/${TAG}_filter_sweep.txt
: > $OUT
run() {
  echo "== $*" >> $OUT
  env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --count-reads 0 --configs= 2>gpurun_out/${TAG}_filter_sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  Q1 %.2f Gq/s (%.3f ms)   Q2 %.2f Gq/s   filter %s' % (d['value']/1e9, d['ms_per_step'], d['extra']['tf23_q2_half_hits']['value']/1e9, json.dumps(d['extra'].get('front_filter'))))" >> $OUT
}
run AIX_INDEX23_FILTER=off
run AIX_BLOOM_BITS=8
run AIX_BLOOM_BITS=8 AIX_FILTER_MINBLOCKS=4
run AIX_BLOOM_BITS=8 AIX_FILTER_MINBLOCKS=5
run AIX_BLOOM_BITS=6
run AIX_BLOOM_BITS=10
run AIX_BLOOM_BITS=12
run AIX_BLOOM_BITS=8 AIX_INDEX23_FILTER=on
cat $OUT
