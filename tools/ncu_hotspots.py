"""Top stall locations of a kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view): samples per instruction,
the dominant stall reason, and a few instructions of context.  Usage: ncu_hotspots.py src.csv [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
body = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
total_inst = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
print(f"{len(body)} SASS instructions, {tot} samples, {total_inst} warp instructions executed")
agg = {}
for h in stall_cols:
    agg[h] = sum(int(r[ci[h]] or 0) for r in body)
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda k: -int(body[k][ci["# Samples"]] or 0))[:top_n]
for k in sorted(order):
    r = body[k]
    n = int(r[ci["# Samples"]] or 0)
    reasons = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"--- #{k} samples {n} ({100.0 * n / tot:.1f}%) exec {r[ci['Instructions Executed']]} :: " +
          ", ".join(f"{nm}={v}" for v, nm in reasons if v))
    for j in range(max(0, k - 2), min(len(body), k + 1)):
        print(f"      {'>>' if j == k else '  '} {body[j][ci['Source']].strip()}")
