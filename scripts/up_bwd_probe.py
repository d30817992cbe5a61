"""Device time of ucd_upsample_bilinear_bwd alone (C ABI called directly, CUDA events, rotating inputs > L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ucd_b200 import _lib
from ucd_b200.losses import ptr, cur_stream, check
shapes = [(24 * 17, 32, 32, 512, 512), (3 * 151, 32, 32, 512, 512), (3 * 20, 32, 64, 512, 1024)]
L = _lib.debug_lib(as_product=True) if os.environ.get('UCD_UPB_UN') or os.environ.get('UCD_UPB_BPS') else _lib.lib()
for planes, h, w, H, W in shapes:
    gs = [torch.randn(planes, H, W, device="cuda") for _ in range(3)]
    gin = torch.empty(planes, h, w, device="cuda")
    def run(k):
        check(L.ucd_upsample_bilinear_bwd(ptr(gs[k % 3]), ptr(gin), planes, h, w, H, W, cur_stream()), "up_bwd")
    for k in range(5): run(k)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for k in range(30): run(k)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    by = planes * H * W * 4
    ref = torch.nn.functional.interpolate(torch.zeros(1, 1, h, w, device="cuda", requires_grad=True), size=(H, W), mode="bilinear")
    print("up_bwd planes=%d %dx%d -> %dx%d: %.1f us  %.0f GB/s" % (planes, H, W, h, w, ms * 1e3, by / ms / 1e6))
