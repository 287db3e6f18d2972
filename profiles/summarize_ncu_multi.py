"""Compact per-kernel table out of one `ncu --set full` report that holds several kernels.
usage: python profiles/summarize_ncu_multi.py <report.ncu-rep> <out.txt>"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    for vals in rows[2:]:
        f.write(f"kernel: {vals[hdr.index('Kernel Name')]}  grid {vals[hdr.index('Grid Size')]}  block {vals[hdr.index('Block Size')]}\n")
        for i, h in enumerate(hdr):
            if h in WANT:
                f.write(f"  {h:76s} {vals[i]:>18s} {units[i]}\n")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        f.write("  stalls per issue: " + ", ".join(f"{h} {v:.2f}" for v, h in sorted(st, reverse=True)[:6]) + "\n\n")
print("wrote", out)
