"""Key counters + warp-stall breakdown of the first kernel in an ncu report (read on the CPU box).
usage: python tools/ncu_summary.py report.ncu-rep [row-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + row]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))
print(d.get("Kernel Name", "")[:100])
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k]:>16s} {u[k]}")
st = []
for h in hdr:
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            st.append((float(d[h]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
tot = sum(v for v, _ in st) or 1.0
print("stall samples:", ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(st, reverse=True)[:10]))
