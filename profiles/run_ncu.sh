#!/usr/bin/env bash
# Profiling recipe (B200_PROFILING.md): the plain run must exit 0 right before each ncu run.
# Usage (on the GPU box, via gpurun): bash profiles/run_ncu.sh <tag> [tests]
set -u
TAG="${1:-r01}"
mkdir -p gpurun_out
if [ "${2:-}" = "tests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
  python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
  python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
fi
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --count-reads 5000000"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tf23_stream -s 3 -c 1 -o gpurun_out/${TAG}_tf23 -f $CMD > gpurun_out/${TAG}_ncu_tf23.log 2>&1
$CMD > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:count13_kernel -s 8 -c 4 -o gpurun_out/${TAG}_count13 -f $CMD > gpurun_out/${TAG}_ncu_count13.log 2>&1
ls -la gpurun_out/
