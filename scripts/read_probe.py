"""HBM read-only ceiling on this box: what a pure streaming READ can reach with different amounts of memory-level
parallelism (register loads / cp.async ring / strided NCHW walk), next to torch.sum and the copy figure.
The read-bound kernels (UNCE / UNKD forward, upsample backward) are judged against the copy bandwidth of
MEASURED_PEAKS.json; this shows how much of the gap is the kernel and how much is the read path itself."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ucd_b200 import _lib

L = _lib.debug_lib()
nbytes = 24 * 17 * 512 * 512 * 4   # the bench workload's logits: 428 MB
buf = torch.randn(nbytes // 4, device="cuda")
us = ctypes.c_float()
print("read probe over %.0f MB (best of 5)" % (nbytes / 1e6))
for mode, name in ((0, "register loads"), (1, "cp.async ring"), (2, "NCHW planes"), (3, "block regions/128thr")):
    for un in (4, 8, 16):
        for bps in ((2, 4, 8) if mode < 3 else (6, 8, 12, 16)):
            if mode == 1 and un == 16 and bps == 8:
                continue  # 8 x 64 KB of ring does not fit
            rc = L.ucd_selftest_read_probe(buf.data_ptr(), nbytes, mode, un, bps, 5, ctypes.byref(us))
            assert rc == 0, rc
            print("%-15s un=%2d blocks/SM=%d  %7.1f us  %6.0f GB/s" % (name, un, bps, us.value, nbytes / us.value / 1e3))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn, by in (("torch.sum", lambda: buf.sum(), nbytes), ("torch copy", lambda: buf.clone(), 2 * nbytes),
                     ("torch fill", lambda: buf.fill_(1.0), nbytes)):
    best = 1e9
    for _ in range(6):
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("%-15s %7.1f us  %6.0f GB/s" % (name, best * 1e3, by / best / 1e6))
