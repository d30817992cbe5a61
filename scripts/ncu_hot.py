"""Top stall sites of one kernel from an `ncu --page source --csv` dump (SASS view).
usage: ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv ; python scripts/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
print(rows[0][1][:80], "| total samples", tot)
agg = {}
for r in body:
    for i in stall_cols:
        v = int(r[i]) if r[i].isdigit() else 0
        agg[hdr[i]] = agg.get(hdr[i], 0) + v
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for idx in sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:n]:
    r = body[idx]
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("%5.1f%%  %s  %-70s %s" % (100 * int(r[si]) / tot, r[0][-5:], r[1][:70], " ".join("%s:%d" % (b, a) for a, b in top if a)))
