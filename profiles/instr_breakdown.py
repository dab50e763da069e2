#!/usr/bin/env python3
"""Executed thread-instructions per unit of work, grouped by source function, for one kernel of an
`ncu --set full --import-source on` report.  The source page of the report gives executed counts per SASS instruction;
`nvdisasm --print-line-info` on the cubin of the same build gives the source line of every SASS instruction; both list
the instructions in address order (the opcode sequences are checked against each other).

  python profiles/instr_breakdown.py <report.ncu-rep> <object.o> <mangled-kernel-prefix> <units> <out.txt>
e.g. profiles/instr_breakdown.py gpurun_out/r02_tf23.ncu-rep aindex_b200/csrc/_obj/tf_query.o \
         _ZN3aix18tf23_stream_kernelILi0ELb1ELi1E 100000000 profiles/r02_tf23_instr.txt
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

GROUPS = [  # (file, first line, last line, label) -- csrc/device_common.cuh, query23.cuh, tf_query.cu
]


def group_of(fn, src_root):
    """label = the enclosing __device__ / __global__ function of (file, line), found by scanning the source upwards"""
    cache = group_of.cache
    f, line = fn
    key = (f, line)
    if key in cache:
        return cache[key]
    path = None
    for root, _, files in os.walk(src_root):
        if f in files:
            path = os.path.join(root, f)
            break
    label = f
    if path:
        lines = open(path, errors="replace").read().split("\n")
        for i in range(min(line, len(lines)) - 1, -1, -1):
            m = re.search(r"(?:__device__|__global__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", lines[i])
            if m and not lines[i].lstrip().startswith("//"):
                label = f"{f}:{m.group(1)}"
                break
    cache[key] = label
    return label


group_of.cache = {}


def main():
    rep, obj, prefix, units, out = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
    src_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "aindex_b200", "csrc")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\." + re.escape(prefix), l))
    end = next(i for i in range(start + 1, len(dis)) if re.match(r"\s*\.section", dis[i]))
    cur, seq = None, []
    for l in dis[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m:
            seq.append((cur, m.group(1).strip()))
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    nxt = next((i for i in range(h + 1, len(rows)) if rows[i] and rows[i][0] in ("Address", "Kernel Name")), len(rows))
    data = [r for r in rows[h + 1:nxt] if len(r) == len(hdr)]
    te, src, sm = hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")

    def op(s):
        t = s.split()
        return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]

    if len(seq) != len(data) or any(op(a[1]) != op(b[src]) for a, b in zip(seq, data)):
        raise SystemExit(f"SASS of {obj} does not match the report ({len(seq)} vs {len(data)} instructions): rebuild or use the matching object")
    per, samples, ops = collections.Counter(), collections.Counter(), collections.Counter()
    for (c, s), r in zip(seq, data):
        g = group_of(c, src_root) if c else "?"
        per[g] += int(r[te])
        samples[g] += int(r[sm])
        ops[op(s)] += int(r[te])
    tot, stot = sum(per.values()), max(1, sum(samples.values()))
    with open(out, "w") as f:
        f.write(f"# {os.path.basename(rep)}: executed thread-instructions per unit ({units:.0f} units), by enclosing source function;\n")
        f.write(f"# stall = share of the warp-stall samples.  total {tot / units:.1f} instructions / unit\n")
        for k, v in per.most_common():
            f.write(f"{v / units:9.1f}  {100 * samples[k] / stot:5.1f}%  {k}\n")
        f.write("# by opcode\n")
        for k, v in ops.most_common(14):
            f.write(f"{v / units:9.1f}  {k}\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
