"""Summarise an ncu report / launch list into a small text table for profiles/.
  ncu_summary.py rep  <file.ncu-rep>       per-kernel metrics of a --set full capture
  ncu_summary.py list <launches.csv>       per-kernel-name totals of a --metrics gpu__time_duration.sum launch list"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

METRICS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
           ("sm__inst_executed_pipe_tensor_subpipe_tcgen05.avg.pct_of_peak_sustained_active", "tcgen05 inst%"),
           ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
           ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu wavefronts%"),
           ("lts__t_bytes.sum", "L2 bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active%"),
           ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("launch__shared_mem_per_block_dynamic", "dyn smem")]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r".*::", "", name)


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== %s  grid %s block %s" % (short(r[idx["Kernel Name"]]), r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for m, label in METRICS:
            if m in idx:
                print("   %-18s %s %s" % (label, r[idx[m]], units[idx[m]]))


def lst(path):
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    tot = OrderedDict()
    for r in rows:
        t = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        t = t / 1e3 if u in ("ns", "nsecond") else (t * 1e3 if u in ("ms", "msecond") else (t * 1e6 if u in ("s", "second") else t))
        k = short(r["Kernel Name"])
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += t
    s = sum(v[1] for v in tot.values())
    print("total %.1f us over %d launches" % (s, len(rows)))
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-44s n=%4d %10.1f us  %5.1f%%" % (k[:44], n, t, 100 * t / s))


if __name__ == "__main__":
    (rep if sys.argv[1] == "rep" else lst)(sys.argv[2])
