"""Per-kernel-name totals of an ncu launch list with gpu__time_duration.sum (+ optional dram__bytes_read/write.sum).
usage: launch_summary.py launches.csv [first_launch] [n_launches]   -> table for profiles/"""
import csv, io, re, sys
from collections import OrderedDict

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
launch = OrderedDict()
for r in rows:
    d = launch.setdefault(int(r["ID"]), {"name": re.sub(r".*::", "", re.sub(r"\(.*", "", r["Kernel Name"])), "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time"):
        d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else (v * 1e6 if u in ("s", "second") else v))
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["rd" if "read" in r["Metric Name"] else "wr"] = v * scale
ids = sorted(launch)
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else len(ids) - first
sel = [launch[i] for i in ids[first:first + n]]
tot = sum(d.get("us", 0.0) for d in sel)
agg = OrderedDict()
for d in sel:
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("rd", 0.0); a[3] += d.get("wr", 0.0)
print("launches %d..%d: total %.1f us over %d launches" % (first, first + len(sel) - 1, tot, len(sel)))
print("%-44s %5s %11s %6s %12s %12s %9s" % ("kernel", "n", "us", "share", "dram rd MB", "dram wr MB", "GB/s"))
for k, (c, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %5d %11.1f %5.1f%% %12.1f %12.1f %9.0f" % (k[:44], c, us, 100 * us / tot, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3 if us else 0))
