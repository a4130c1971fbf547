#!/usr/bin/env python
"""Per-function instruction / stall-sample shares of one kernel from an .ncu-rep with source info.

    python tools/func_breakdown.py <rep.ncu-rep> <libnnuepack.so> <kernel> <units per warp>

Maps ncu's per-SASS-instruction counters to source lines (tools/sass_lines.py) and sums them by the
device function the line belongs to (function spans are read from the .cuh/.cu files)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nnue_data_compress_b200", "csrc")


def function_spans(path):
    spans, cur, depth, start = [], None, 0, 0
    for i, line in enumerate(open(path, errors="ignore"), 1):
        if re.match(r"\s*namespace\b.*\{\s*$", line) or re.match(r"\s*\}\s*//\s*namespace", line) or 'extern "C"' in line:
            continue  # namespace braces do not nest functions
        if depth == 0:
            m = re.search(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$", line)
            if m and not line.strip().startswith(("//", "#", "*")) and m.group(1) not in ("if", "for", "while", "switch", "return"):
                cur, start = m.group(1), i
        depth += line.count("{") - line.count("}")
        if depth == 0 and cur and "}" in line:
            spans.append((start, i, cur))
            cur = None
    return spans


def main():
    rep, so, kernel, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], stdout=f,
                       stderr=subprocess.DEVNULL)
        page = f.name
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_lines.py"), so, kernel, page, "100000"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    spans = {}
    tot = {}
    total_inst = total_smp = 0.0
    for line in out.splitlines():
        m = re.match(r"\s*([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\S+):(\d+)", line)
        if not m:
            continue
        inst, smp, fname, ln = float(m.group(1)), float(m.group(3)), m.group(4), int(m.group(5))
        path = os.path.join(CSRC, fname)
        name = fname
        if os.path.exists(path):
            if fname not in spans:
                spans[fname] = function_spans(path)
            for a, b, fn in spans[fname]:
                if a <= ln <= b:
                    name = f"{fname}:{fn}"
                    break
        t = tot.setdefault(name, [0.0, 0.0])
        t[0] += inst
        t[1] += smp
        total_inst += inst
        total_smp += smp
    print(f"kernel {kernel}: {total_inst / units:.0f} warp instructions per unit ({units:g} units per warp)")
    print(f"{'inst/unit':>10} {'%inst':>6} {'%stall samples':>14}  function")
    for name, (inst, smp) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
        print(f"{inst / units:10.1f} {100 * inst / total_inst:6.1f} {smp:14.1f}  {name}")


if __name__ == "__main__":
    main()
