"""Tensor-pipe time per column tile for the sweep-1 MMA mix under different tensor-memory placements (probe)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ucd_b200 import _lib
L = _lib.debug_lib()
torch.zeros(1, device="cuda")
# (label, s_ts, s_a, s_acc0, s_acc1, v_n256, v_a, v_acc)
cfgs = [
    ("SS S@0        | V N=128x2 A@128 acc@256 (kernel layout)", 0, 0, 0, 0, 0, 128, 256),
    ("SS S@0/S@128  | V N=128x2 A@0   acc@256 (E over S, two S buffers)", 0, 0, 0, 128, 0, 0, 256),
    ("SS S@0        | V N=128x2 A@0   acc@256", 0, 0, 0, 0, 0, 0, 256),
    ("SS S@0/S@128  | V N=256   A@0   acc@256 (round-1 layout)", 0, 0, 0, 128, 1, 0, 256),
    ("SS S@0        | V N=256   A@128 acc@256", 0, 0, 0, 0, 1, 128, 256),
    ("SS S@0        | V N=128x2 A@192 acc@256", 0, 0, 0, 0, 0, 192, 256),
    ("SS S@384      | V N=128x2 A@256 acc@0", 0, 0, 384, 384, 0, 256, 0),
    ("TS A@0 S@128  | V N=128x2 A@192 acc@256", 1, 0, 128, 128, 0, 192, 256),
    ("TS A@0 S@128/S@192?no: S@128 | V N=256 A@192 acc@256", 1, 0, 128, 128, 1, 192, 256),
]
for c in cfgs:
    out = ctypes.c_float()
    _lib.check(L.ucd_selftest_mma_mix(*c[1:], 512, ctypes.byref(out)), "mma_mix")
    print("%-75s %7.0f clk / tile" % (c[0], out.value))
