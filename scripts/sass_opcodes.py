#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libucd_b200.so (cuobjdump -sass): the evidence that the sweeps are
tcgen05 / TMEM / bulk-copy native (UTCHMMA, LDTM / STTM, UBLKCP, UTCBAR) and that no legacy tensor path (HMMA / HGMMA)
or wrapper loop (ELECT + BRA.U.ANY around every uniform-datapath instruction) is left.  No GPU needed.
    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ucd_b200", "libucd_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ELECT", "BRA.U.ANY",
       "MUFU.EX2", "MUFU.LG2", "MUFU.RCP", "F2FP", "FMNMX3", "HMMA", "HGMMA", "LDG", "STG", "LDS", "STS", "BAR", "LDGSTS"]
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        for k in KEY:
            if op == k or op.startswith(k + "."):
                hist[kern][k] += 1
print("# SASS opcode histogram of %s (sm_100a)\n" % os.path.basename(lib))
print("| kernel | instr | " + " | ".join(KEY) + " |")
print("|---|---:|" + "---:|" * len(KEY))
tot = collections.Counter()
for k, h in hist.items():
    print("| `%s` | %d | " % (k.replace("void ", "")[:60], h["_total"]) + " | ".join(str(h[x]) if h[x] else "" for x in KEY) + " |")
    tot.update(h)
print("| **all kernels** | %d | " % tot["_total"] + " | ".join(str(tot[x]) for x in KEY) + " |")
print("\ntcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, "
      "mbarrier -> SYNCS; HMMA/HGMMA (legacy mma.sync / wgmma) must be 0.  BRA.U.ANY counts the ELECT wrapper loops "
      "ptxas puts around uniform-datapath instructions issued under `if (lane == 0)`: 0 in the sweeps since round 2.")
