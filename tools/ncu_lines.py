"""Per-CUDA-line summary of an `ncu --page source --csv --print-source cuda,sass` dump (samples, instructions, shared-memory
wavefronts).  usage: python tools/ncu_lines.py dump.csv <kernel-name-substring> [top]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    key = sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    idx = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
    sel, started = [], False
    for k, i in enumerate(idx):
        hit = key in rows[i][1]
        if hit:
            started = True
        if started and not hit:
            break
        if hit:
            end = idx[k + 1] if k + 1 < len(idx) else len(rows)
            fp = rows[i - 1][1] if rows[i - 1][0] == "File Path" else "?"
            hdr = rows[i + 1]
            for r in rows[i + 2:end]:
                if r and r[0] not in ("", "File Path", "Function Name", "Line No"):
                    sel.append((fp.split("/")[-1], r, hdr))
    h = sel[0][2]
    iw, iwe = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Excessive")
    ts = sum(num(r[4]) for _, r, _ in sel)
    ti = sum(num(r[7]) for _, r, _ in sel)
    tw = sum(num(r[iw]) for _, r, _ in sel)
    print("samples %d  warp instructions %d  shared wavefronts %d (excessive %d)" % (ts, ti, tw, sum(num(r[iwe]) for _, r, _ in sel)))
    sel.sort(key=lambda t: -num(t[1][4]))
    for fp, r, _ in sel[:top]:
        print("%-16s L%-5s samp %5.1f%% inst %5.1f%% wav %5.1f%% exc %9s | %s" % (
            fp, r[0], 100 * num(r[4]) / max(ts, 1), 100 * num(r[7]) / max(ti, 1), 100 * num(r[iw]) / max(tw, 1), r[iwe], r[1].strip()[:100]))


main()
