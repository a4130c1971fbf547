"""CPU: one line per kernel from the CSV of tools/ncu_counts.sh: python tools/counts_summary.py gpurun_out/counts_main.csv [positions]"""
import csv
import sys

n = float(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
d = {}
for r in rows[1:]:
    d.setdefault(r[ik].split("(")[0], {})[r[im]] = float(r[iv].replace(",", ""))
for k, m in d.items():
    g = m.get
    print(f"{k}: {g('gpu__time_duration.sum') / 1e6:.3f} ms | inst/pos {g('smsp__inst_executed.sum') / n * 32:.0f} "
          f"(alu {g('sm__inst_executed_pipe_alu.sum') / n * 32:.0f} fma {g('sm__inst_executed_pipe_fma.sum') / n * 32:.0f} "
          f"lsu {g('sm__inst_executed_pipe_lsu.sum') / n * 32:.0f}) lanes {g('smsp__thread_inst_executed.sum') / g('smsp__inst_executed.sum'):.1f} | "
          f"issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f}% alu {g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.0f}% "
          f"warps {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f}% long_sb {g('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'):.2f} | "
          f"dram rd {g('dram__bytes_read.sum') / n:.1f} wr {g('dram__bytes_write.sum') / n:.1f} B/pos, L1 {g('l1tex__t_sector_hit_rate.pct'):.0f}% L2 {g('lts__t_sector_hit_rate.pct'):.0f}%")
