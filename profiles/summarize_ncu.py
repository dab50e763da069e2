#!/usr/bin/env python3
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/, scratch) into the small text /
JSON summaries that are committed under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/<tag>_launches.csv  profiles/<name>.txt
  python profiles/summarize_ncu.py kernel   gpurun_out/<tag>_<k>.ncu-rep   profiles/<name>.txt [traffic-key [units-profiled]]

`kernel` also updates profiles/traffic.json[traffic-key] = DRAM read+write bytes per launch,
which bench.py reports as roofline.traffic (B200_PROFILING.md: from one `ncu --set full` capture).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",  # address-divergence unit: MATCH.ANY, indexed constant loads, barriers
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",  # register spills
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]

UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki]
        v = float(r[vi].replace(",", ""))
        if r[ui] == "us":
            v *= 1e3
        elif r[ui] == "ms":
            v *= 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "aix::" in k)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({os.path.basename(src)}): gpu__time_duration.sum per kernel, cold-cache and serialised\n")
        f.write(f"# {len(rows) - 1} launches, {tot / 1e6:.3f} ms total, {ours / 1e6:.3f} ms in aix:: kernels; the rest is torch\n")
        f.write("# synthetic-data generation and CUB sorts of the (untimed) index setup\n")
        f.write(f"{'launches':>8} {'total_ms':>10} {'avg_ms':>9} {'share':>6}  kernel\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{a[0]:8d} {a[1] / 1e6:10.3f} {a[1] / a[0] / 1e6:9.4f} {100 * a[1] / tot:5.1f}%  {k[:150]}\n")
    print(open(dst).read()[:3000])


def kernel(src, dst, traffic_key=None, units_profiled=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({os.path.basename(src)}), one block per captured launch\n")
        traffic = []
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')][:160]}\n")
            vals = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    vals[k] = (r[i], units[i])
                    f.write(f"{k:90s} {r[i]:>18s} {units[i]}\n")
            try:
                rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * UNIT_SCALE[vals["dram__bytes_read.sum"][1]]
                wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * UNIT_SCALE[vals["dram__bytes_write.sum"][1]]
                traffic.append(rd + wr)
                f.write(f"{'dram traffic (read + write), bytes per launch':90s} {rd + wr:18.0f} byte\n")
            except Exception:
                pass
    if traffic_key and traffic:
        p = os.path.join(HERE, "traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[traffic_key] = {"bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic),
                          "source": os.path.relpath(dst, os.path.dirname(HERE))}
        if units_profiled:
            d[traffic_key]["units_profiled"] = int(units_profiled)  # bench.py scales the traffic to the units of its own launch
        json.dump(d, open(p, "w"), indent=1, sort_keys=True)
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else None)
