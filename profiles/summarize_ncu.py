"""Turns one `ncu --set full --import-source on` report into the text summaries kept under profiles/.
usage: python profiles/summarize_ncu.py <report.ncu-rep> <out_prefix>
writes <out_prefix>_metrics.txt (selected raw metrics + stall reasons) and <out_prefix>_hot_lines.txt (source lines
of lcsim_core.cuh / lcsim_b200.cu ranked by stall samples, with executed instructions and threads per instruction)."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
with open(out + "_metrics.txt", "w") as f:
    f.write(f"kernel: {vals[hdr.index('Kernel Name')]}  grid {vals[hdr.index('Grid Size')]}  block {vals[hdr.index('Block Size')]}\n")
    for i, h in enumerate(hdr):
        if h in WANT:
            f.write(f"{h:78s} {vals[i]:>18s} {units[i]}\n")
    f.write("--- warps stalled per issue-active cycle, by reason\n")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            st.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    for v, h in sorted(st, reverse=True)[:10]:
        f.write(f"{h:32s} {v:6.2f}\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines, fname, cols = [], None, None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname, cols = r[1].split("/")[-1], None
        continue
    if r and r[0] == "Line No":
        cols = r
        continue
    if cols and fname and len(r) == len(cols) and r[2] == "-":  # per-source-line aggregate row (SASS rows carry an address)
        try:
            samp, inst, thr = int(r[cols.index("# Samples")] or 0), int(r[cols.index("Instructions Executed")] or 0), int(r[cols.index("Thread Instructions Executed")] or 0)
        except ValueError:
            continue
        if inst or samp:
            lines.append((samp, inst, thr, fname, int(r[0]), r[1].strip()))
tot_s, tot_i = sum(x[0] for x in lines) or 1, sum(x[1] for x in lines) or 1
with open(out + "_hot_lines.txt", "w") as f:
    f.write(f"total inst {tot_i} samples {tot_s}\n")
    for samp, inst, thr, fn, ln, txt in sorted(lines, reverse=True)[:40]:
        f.write(f"{fn[:10]:10s} {ln:5d} samp {100 * samp / tot_s:5.1f}%  inst {inst:9d} ({100 * inst / tot_i:4.1f}%) thr/inst {thr / max(inst, 1):5.1f}  {txt[:110]}\n")
    f.write("--- by executed instructions\n")
    for samp, inst, thr, fn, ln, txt in sorted(lines, key=lambda x: -x[1])[:50]:
        f.write(f"{fn[:10]:10s} {ln:5d} samp {100 * samp / tot_s:5.1f}%  inst {inst:9d} ({100 * inst / tot_i:4.1f}%) thr/inst {thr / max(inst, 1):5.1f}  {txt[:110]}\n")
    # instructions / stall samples per 25-line bucket of each file: a coarse per-function picture
    f.write("--- by 25-line bucket\n")
    buckets = {}
    for samp, inst, thr, fn, ln, txt in lines:
        k = (fn, ln // 25 * 25)
        b = buckets.setdefault(k, [0, 0])
        b[0] += samp; b[1] += inst
    for (fn, ln), (samp, inst) in sorted(buckets.items(), key=lambda kv: -kv[1][1])[:40]:
        f.write(f"{fn[:14]:14s} {ln:5d}-{ln + 24:5d} samp {100 * samp / tot_s:5.1f}%  inst {100 * inst / tot_i:5.1f}%\n")
print("wrote", out + "_metrics.txt", out + "_hot_lines.txt")
