import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ucd_b200 import _lib
L = _lib.lib()
names = ["SS N=64", "SS N=128", "SS N=256 (B MN-major)", "TS N=64", "TS N=128", "TS N=256 (B MN-major)", "SS N=256 (B K-major)",
         "TS N=128 alternating 2 accumulators", "SS N=128 alternating 2 accumulators", "2 threads: SS N=128", "2 threads: TS N=128", "2 threads: SS N=64", "elect_one_sync: SS N=128", "elect_one_sync: TS N=128", "elect_one_sync: SS N=64", "elect_one_sync: TS N=256 (B MN-major)",
         "elect_one_sync, 2 warps: 16 SS N=128 + 8 TS N=256 per tile (cycles per tile / 16)"]
torch.zeros(1, device="cuda")
for mode, nm in enumerate(names):
    out = ctypes.c_float()
    _lib.check(L.ucd_selftest_mma_rate(mode, 4096, ctypes.byref(out)), "mma_rate")
    print("%-40s %7.1f cycles / tcgen05.mma" % (nm, out.value))
