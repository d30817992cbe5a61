import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ucd_b200 import _lib
L = _lib.debug_lib()
names = ["SS N=64", "SS N=128", "SS N=256 (B MN-major)", "TS N=64", "TS N=128", "TS N=256 (B MN-major)", "SS N=256 (B K-major)",
         "TS N=128 alternating 2 accumulators", "SS N=128 alternating 2 accumulators", "2 threads: SS N=128", "2 threads: TS N=128", "2 threads: SS N=64", "elect_one_sync: SS N=128", "elect_one_sync: TS N=128", "elect_one_sync: SS N=64", "elect_one_sync: TS N=256 (B MN-major)",
         "2 warps, per tile 16 S (SS N=128) + 8 V (TS N=256): cycles per tile / 16",
         "2 warps, per tile 16 S (SS N=128) + 16 V (TS N=128): cycles per tile / 16",
         "2 warps, per tile 16 S (TS N=128, anchors in TMEM) + 16 V (TS N=128): cycles per tile / 16",
         "2 warps, per tile 8 S TS + 8 S SS + 16 V (TS N=128): cycles per tile / 16"]
torch.zeros(1, device="cuda")
for mode, nm in enumerate(names):
    out = ctypes.c_float()
    _lib.check(L.ucd_selftest_mma_rate(mode, 4096, ctypes.byref(out)), "mma_rate")
    print("%-90s %7.1f cycles / tcgen05.mma" % (nm, out.value))
