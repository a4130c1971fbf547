#!/usr/bin/env python
"""Where do the idle lanes go? Per source function: warp instructions, active lanes per instruction and
the share of the kernel's lost lane-slots (32 - active lanes, weighted by execution count).

    ncu -i X.ncu-rep --page source --csv --kernel-name regex:<kernel> > page.csv
    python tools/lane_waste.py <libnnuepack.so> <kernel> page.csv
"""
import csv
import importlib.util
import os
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("sass_lines", os.path.join(HERE, "sass_lines.py"))
sl = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sl)
spec = importlib.util.spec_from_file_location("func_breakdown", os.path.join(HERE, "func_breakdown.py"))
fb = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fb)


def main():
    so, kernel, page = sys.argv[1:4]
    ins = sl.disasm_lines(so, kernel)
    rows = list(csv.reader(open(page)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[h]
    ci, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = [r for r in rows[h + 1:] if len(r) > ci and r[ci].isdigit()]
    spans = {}
    per = defaultdict(lambda: [0, 0])
    tot = [0, 0]
    for (addr, text, loc), r in zip(ins, data):
        fname, ln = loc
        name = fname
        path = os.path.join(fb.CSRC, fname)
        if os.path.exists(path):
            if fname not in spans:
                spans[fname] = fb.function_spans(path)
            for a, b, fn in spans[fname]:
                if a <= ln <= b:
                    name = f"{fname}:{fn}"
                    break
        w, t = int(r[ci]), int(r[ti])
        per[name][0] += w
        per[name][1] += t
        tot[0] += w
        tot[1] += t
    lost_total = 32 * tot[0] - tot[1]
    print(f"{kernel}: {tot[1] / tot[0]:.1f} active lanes per instruction; lost lane-slots by function")
    print(f"{'%inst':>6} {'lanes':>6} {'%lost':>6}  function")
    for name, (w, t) in sorted(per.items(), key=lambda kv: -(32 * kv[1][0] - kv[1][1]))[:24]:
        print(f"{100 * w / tot[0]:6.1f} {t / max(w, 1):6.1f} {100 * (32 * w - t) / lost_total:6.1f}  {name}")


if __name__ == "__main__":
    main()
