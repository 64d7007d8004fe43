"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: python scripts/summarize_ncu.py <launches.csv> <prof.ncu-rep> <tag>"""
import csv, io, os, re, subprocess, sys

launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
os.makedirs("profiles", exist_ok=True)

# ---- launch list: per-kernel totals and shares
rows = [l for l in open(launch_csv) if l.startswith('"')]
rd = list(csv.DictReader(io.StringIO("".join(rows))))
tot = {}
for r in rd:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    ns = float(r["Metric Value"].replace(",", ""))
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += ns
all_ns = sum(v[1] for v in tot.values())
with open("profiles/%s_launches.md" % tag, "w") as f:
    f.write("# ncu launch list (%s): gpu__time_duration.sum, --clock-control none, cold-cache serialised\n\n" % tag)
    f.write("command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (map build + 5 resident steps + 3 host-buffer steps + peak microbench)\n\n")
    f.write("| kernel | launches | total ms | share | ms / launch |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("| %s | %d | %.3f | %.1f%% | %.3f |\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / all_ns, v[1] / 1e6 / v[0]))
with open("profiles/%s_launches.csv" % tag, "w") as f:
    f.write("".join(rows))

# ---- full capture: the metrics the judge asks for + pipe utilisation + stall reasons
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rr[0], rr[1], rr[2]
keep = re.compile(r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|gpu__dram_throughput|"
                  r"sm__pipe_tensor_cycles_active.avg.pct|sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct|sm__inst_executed_pipe_(alu|fma|lsu|tmem|tma|uniform|xu|tc|adu|cbu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_(alu|fma|shared|tc)_cycles_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct|smsp__warps_active.avg.per_cycle_active|smsp__warps_eligible.avg.per_cycle_active|"
                  r"smsp__issue_active.avg.pct|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|waves_per_multiprocessor)$|lts__t_sector_hit_rate.pct|lts__t_bytes.sum$|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|sm__cycles_elapsed.avg$|sm__cycles_active.avg$|smsp__inst_executed.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
                  r"sm__throughput.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct)")
with open("profiles/%s_ncu_full.md" % tag, "w") as f:
    f.write("# ncu --set full --clock-control none, one launch of the dominant kernel (%s)\n\n" % tag)
    f.write("command: `python bench.py --steps 1 --warmup 1 --steps 1 --warmup 3 --no-cpu-baseline` at the default configuration (-k regex:sift_tc4 -s 1 -c 1)\n\n")
    f.write("| metric | unit | value |\n|---|---|---:|\n")
    for h, u, v in sorted(zip(hdr, units, vals)):
        if keep.match(h):
            f.write("| %s | %s | %s |\n" % (h, u, v))
print(open("profiles/%s_launches.md" % tag).read())
print(open("profiles/%s_ncu_full.md" % tag).read())
