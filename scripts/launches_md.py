"""ncu launch list (gpu__time_duration.sum CSV) -> per-kernel totals and shares.  usage: python scripts/launches_md.py <launches.csv> <tag> "<command>" """
import csv, io, re, sys
launch_csv, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [l for l in open(launch_csv) if l.startswith('"')]
rd = list(csv.DictReader(io.StringIO("".join(rows))))
tot = {}
for r in rd:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    ns = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit", "ns") in ("us", "usecond"):
        ns *= 1e3
    elif r.get("Metric Unit", "ns") in ("ms", "msecond"):
        ns *= 1e6
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += ns
all_ns = sum(v[1] for v in tot.values())
with open("profiles/%s_launches.md" % tag, "w") as f:
    f.write("# ncu launch list (%s): gpu__time_duration.sum, --clock-control none, cold-cache serialised\n\n" % tag)
    f.write("command: `%s` (first %d launches)\n\n" % (cmd, len(rd)))
    f.write("| kernel | launches | total ms | share | ms / launch |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("| %s | %d | %.3f | %.1f%% | %.3f |\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / all_ns, v[1] / 1e6 / v[0]))
open("profiles/%s_launches.csv" % tag, "w").write("".join(rows))
print(open("profiles/%s_launches.md" % tag).read())
