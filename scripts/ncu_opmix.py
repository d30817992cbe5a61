"""Dynamic opcode mix of one kernel from an `ncu --page source --csv` dump: warp instructions executed per opcode and
per processed element.  usage: python scripts/ncu_opmix.py src.csv <elements>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_elem = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ie, src = hdr.index("Instructions Executed"), hdr.index("Source")
body = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body)
print(rows[0][1][:90])
print("warp instructions executed %d | thread instructions per element %.2f" % (tot, tot * 32 / n_elem))
h = collections.Counter()
for r in body:
    t = r[src].split()
    op = t[1] if t[0].startswith("@") else t[0]
    h[op.split(".")[0].rstrip(";")] += int(r[ie])
for k, v in h.most_common(22):
    print("  %-10s %5.1f%%  %.2f per element" % (k, 100 * v / tot, v * 32 / n_elem))
