"""Brief markdown summary of one kernel from an .ncu-rep: python scripts/ncu_brief.py <rep> <out.md> <title>"""
import csv, io, re, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rr[0], rr[1], rr[2]
keep = re.compile(r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__bytes_read.sum.pct_of_peak_sustained_elapsed|"
                  r"sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|tmem|tma|uniform|xu|tc|adu|cbu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_(alu|fma|fmaheavy|shared|tc|tensor|xu)_cycles_active.avg.pct_of_peak_sustained_active|"
                  r"sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct_of_peak_sustained_active|"
                  r"smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum$|launch__(registers_per_thread|grid_size|block_size|cluster_dim_x|waves_per_multiprocessor)$|"
                  r"lts__t_sector_hit_rate.pct|lts__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|"
                  r"smsp__inst_executed_op_.*popc.*|sm__sass_thread_inst_executed_op_integer_pred_on.sum$)")
stall = sorted(((h, v) for h, v in zip(hdr, vals) if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")),
               key=lambda t: -float(t[1] or 0))[:6]
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none: %s\n\n| metric | unit | value |\n|---|---|---:|\n" % title)
    for h, u, v in sorted(zip(hdr, units, vals)):
        if keep.match(h):
            f.write("| %s | %s | %s |\n" % (h, u, v))
    f.write("\n| top warp stall reasons (per issue-active cycle) | ratio |\n|---|---:|\n")
    for h, v in stall:
        f.write("| %s | %s |\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
print(open(out).read())
