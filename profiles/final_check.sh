mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r05a_pytest.log 2>&1; echo rc=$? >> gpurun_out/r05a_pytest.log)
(python __graft_entry__.py --smoke > gpurun_out/r05a_smoke.log 2>&1; echo rc=$? >> gpurun_out/r05a_smoke.log)
(timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r05a_bench_ref.json 2> gpurun_out/r05a_bench_ref.err)
(timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r05a_bench.json 2> gpurun_out/r05a_bench.err; echo rc=$? >> gpurun_out/r05a_bench.err)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --count-reads 5000000 --configs="
$CMD > gpurun_out/r05a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r05a_launches.csv $CMD > gpurun_out/r05a_ncu_launches.log 2>&1
python profiles/summarize_ncu.py launches gpurun_out/r05a_launches.csv gpurun_out/r05a_launches.txt > /dev/null 2>&1
rm -f gpurun_out/r05a_launches.csv
