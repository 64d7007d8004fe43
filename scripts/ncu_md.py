"""ncu report -> the tracked markdown summary under profiles/.  usage: python scripts/ncu_md.py <prof.ncu-rep> <tag> "<command line captured>" """
import csv, io, re, subprocess, sys
rep, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
keep = re.compile(r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|gpu__dram_throughput|"
                  r"sm__pipe_tensor_cycles_active.avg.pct|sm__pipe_tensor_subpipe_.*cycles_active.avg.pct|sm__inst_executed_pipe_(alu|fma|fp64|lsu|tmem|tma|uniform|xu|tc|adu|cbu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_(alu|fma|fp64|shared|tc)_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct|smsp__warps_active.avg.per_cycle_active|smsp__warps_eligible.avg.per_cycle_active|"
                  r"smsp__issue_active.avg.pct|launch__(registers_per_thread|grid_size|block_size|cluster_size|shared_mem_per_block_dynamic|waves_per_multiprocessor)$|lts__t_sector_hit_rate.pct|lts__t_bytes.sum$|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|sm__cycles_elapsed.avg$|sm__cycles_active.avg$|smsp__inst_executed.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
                  r"sm__throughput.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct|lts__t_sectors_srcunit_tex_op_read.sum$|l1tex__m_xbar2l1tex_read_bytes.sum$)")
out = ["# ncu --set full --clock-control none (%s)\n" % tag, "command: `%s`\n" % cmd]
for vals in rr[2:]:
    if len(vals) != len(hdr):
        continue
    d = dict(zip(hdr, vals))
    out.append("\n## %s  (grid %s, block %s)\n\n| metric | unit | value |\n|---|---|---:|" % (d.get("Kernel Name", "?")[:120], d.get("Grid Size", "?"), d.get("Block Size", "?")))
    for h, u, v in sorted(zip(hdr, units, vals)):
        if keep.match(h):
            out.append("| %s | %s | %s |" % (h, u, v))
open("profiles/%s.md" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))
