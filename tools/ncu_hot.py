"""Hot SASS instructions of an `ncu --page source --csv` export: python tools/ncu_hot.py file.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[h]
si = hdr.index('# Samples'); src = hdr.index('Source'); ie = hdr.index('Instructions Executed')
data = []
for i, r in enumerate(rows[h + 1:]):
    if len(r) <= si or r[0] == 'Address':
        continue
    try:
        data.append((int(r[si] or 0), r[src].strip(), int(r[ie] or 0), i))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2
print('total samples', tot, 'instructions', len(data))
for d in data:
    if d[0] > tot * pct / 100:
        print(f"{d[3]:5d} {100*d[0]/tot:5.1f}% exec {d[2]:9d}  {d[1][:100]}")
