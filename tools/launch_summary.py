"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel count, total, average, share.  Usage: launch_summary.py csv [out.md]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r
        continue
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1e-3)
    name = d['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    if '--by-grid' in sys.argv:
        name += ' grid=' + d.get('Grid Size', '?').replace(' ', '')
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out = ["| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1]/v[0]:.1f} | {100*v[1]/tot:.1f}% |")
txt = "\n".join(out)
if len(sys.argv) > 2 and not sys.argv[2].startswith('--'):
    open(sys.argv[2], 'w').write(txt + "\n")
print(txt)
