for mb in 1 5 6 1 5; do
  AIX_TF23_MINBLOCKS=$mb python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --count-reads 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('minblocks $mb: Q1 %.2f Gq/s (%.3f ms)   Q2 %.2f Gq/s' % (d['value']/1e9, d['ms_per_step'], d['extra']['tf23_q2_half_hits']['value']/1e9))"
done
