"""Scratch: per-launch table from an ncu --metrics gpu__time_duration.sum csv. usage: launch_table.py file.csv [skip] [count]"""
import csv, io, re, sys
txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cnt = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
tot = 0
for i, r in enumerate(rows[skip:skip + cnt]):
    t = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    t = t / 1e3 if u == 'ns' else (t * 1e3 if u == 'ms' else t)
    tot += t
    print(i, re.sub(r'.*::', '', re.sub(r'\(.*', '', r['Kernel Name']))[:28], r['Grid Size'], '%.1f' % t)
print('total us', tot)
