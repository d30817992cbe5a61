import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ucd_b200 import _lib
L = _lib.debug_lib()
names = ["8 ex2", "4 cvt.bf16x2 + 8 fadd", "8 ex2 + 4 cvt.bf16x2", "8 ex2 + int round pack (8 iadd, 4 prmt)", "8 fadd"]
torch.zeros(1, device="cuda")
for warps in (4, 8, 16):
    for mode, nm in enumerate(names):
        out = ctypes.c_float()
        _lib.check(L.ucd_selftest_pipe_rate(mode, warps, 4096, ctypes.byref(out)), "pipe_rate")
        print("warps %2d  %-42s %7.1f cycles / iteration (8 values per thread)" % (warps, nm, out.value))
