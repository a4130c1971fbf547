#!/usr/bin/env python
"""Attributes ncu per-instruction counters to CUDA source lines.

ncu's CSV export of the source page carries metrics only in SASS view, and nvdisasm knows the
line of every SASS instruction; both list a function's instructions in the same order.

    python tools/sass_lines.py <libnnuepack.so> <kernel-substring> <ncu_sass_page.csv> [top]

where the CSV comes from
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:<kernel> > ncu_sass_page.csv
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def disasm_lines(so, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                             text=True).stdout
        infn = False
        cur = ("?", 0)
        for line in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                infn = kernel in m.group(1)
                continue
            if not infn:
                continue
            if line.startswith("\t.section") or line.strip().startswith(".section"):
                infn = False
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip(), cur))
        if out:
            break
    return out


def main():
    so, kernel, page = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    ins = disasm_lines(so, kernel)
    rows = list(csv.reader(open(page)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[h]
    ci, si, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    ti = hdr.index("Thread Instructions Executed")
    data = [r for r in rows[h + 1:] if len(r) > ci and r[ci].isdigit()]
    if len(data) != len(ins):
        print(f"warning: {len(data)} ncu instructions vs {len(ins)} disassembled", file=sys.stderr)
    per = defaultdict(lambda: [0, 0, 0])
    total = [0, 0, 0]
    for (addr, text, loc), r in zip(ins, data):
        vals = (int(r[ci]), int(r[ti]), int(r[smp] or 0))
        for k in range(3):
            per[loc][k] += vals[k]
            total[k] += vals[k]
    warps = max(int(data[0][ci]), 1)
    print(f"total warp-instructions {total[0]}  thread-instructions {total[1]}  samples {total[2]}  warps {warps}")
    print(f"{'inst/warp':>10} {'%inst':>6} {'%smpl':>6}  location")
    for loc, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0] / warps:10.1f} {100 * v[0] / total[0]:6.1f} {100 * v[2] / max(total[2], 1):6.1f}  {loc[0]}:{loc[1]}")


if __name__ == "__main__":
    main()
