#!/usr/bin/env python
"""Prints the per-launch durations of an `ncu --metrics gpu__time_duration.sum --csv` log.

    python tools/launch_summary.py gpurun_out/launches.csv [last_n]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = []
    for r in rows[h + 1:]:
        if len(r) > vi and r[vi].replace(".", "").replace(",", "").isdigit():
            v = float(r[vi].replace(",", ""))
            unit = r[ui]
            us = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
            out.append((r[ki].split("(")[0].replace("void ", ""), us))
    last = int(sys.argv[2]) if len(sys.argv) > 2 else len(out)
    for name, us in out[-last:]:
        print(f"{us:12.1f} us  {name}")


if __name__ == "__main__":
    main()
