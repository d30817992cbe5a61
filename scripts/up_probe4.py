import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ucd_b200 as U
dev = "cuda"
lr = torch.randn(24, 17, 32, 32, device=dev)
big = [torch.randn(24, 17, 512, 512, device=dev) for _ in range(3)]
dst = [torch.empty(24, 17, 512, 512, device=dev) for _ in range(3)]
def ev(): return torch.cuda.Event(enable_timing=True)
def measure(name, pre, op):
    ts = []
    for k in range(26):
        pre(k)
        a, b = ev(), ev()
        a.record(); op(k); b.record()
        ts.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ts[6:])
    print("%-60s median %.1f us  min %.1f us" % (name, 1e3 * t[len(t) // 2], 1e3 * t[0]))
pres = {"alone": lambda k: None, "after sum (read)": lambda k: big[k % 3].sum(), "after mul_ (r+w)": lambda k: big[k % 3].mul_(1.0001)}
ops = {"torch fill_ 428 MB": lambda k: dst[k % 3].fill_(1.0),
       "torch zero_ (memset) 428 MB": lambda k: dst[k % 3].zero_(),
       "ucd upsample fwd 428 MB": lambda k: U.interpolate_bilinear(lr, (512, 512)),
       "torch copy_ 428 MB -> 428 MB": lambda k: dst[k % 3].copy_(big[(k + 1) % 3])}
for on, op in ops.items():
    for pn, pre in pres.items():
        measure(on + " | " + pn, pre, op)
