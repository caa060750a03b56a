"""Basic-block level instruction/sampling shares from an ncu report (source page, SASS).  Usage: sass_segments.py rep kernel-regex"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]; ci = h.index('Instructions Executed'); cs = h.index('Source'); csm = h.index('# Samples')
data = [(r[cs].strip(), int(r[ci]), int(r[csm])) for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print('total inst %.1fM  samples %d  sass lines %d' % (tot / 1e6, ts, len(data)))
segs = []; start = 0; prev = None
for i, (s, n, sm) in enumerate(data):
    if prev is None or abs(n - prev) > 0.02 * max(n, prev, 1):
        if prev is not None: segs.append((start, i - 1, prev))
        start = i; prev = n
segs.append((start, len(data) - 1, prev))
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
for a, b, n in segs:
    t = sum(d[1] for d in data[a:b + 1]); smp = sum(d[2] for d in data[a:b + 1])
    if t > thr * tot or smp > thr * ts:
        ops = {}
        for d in data[a:b + 1]:
            parts = d[0].split()
            op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
        print(f"sass[{a:4d}-{b:4d}] len={b-a+1:4d} exec/inst={n/1e6:6.2f}M inst={100*t/tot:5.1f}% samples={100*smp/ts:5.1f}% {top}")
