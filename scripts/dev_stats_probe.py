"""gated_stats_multi vs one gated_stats pass per consumer on the decoder's skip sources (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K
N = 48
for name, hw, C, gated, ncons in [("f1", 512, 64, False, 4), ("x_3_3", 512, 64, True, 3), ("x_2_3", 512, 64, True, 2),
                                  ("f2", 256, 256, False, 3), ("x_2_2", 256, 256, True, 2)]:
    x = torch.randn(N, hw, hw, C, device="cuda").bfloat16()
    cg = torch.rand(N, C, device="cuda") if gated else None
    sg = torch.rand(N, hw, hw, device="cuda") if gated else None
    cons = [(torch.randn(C, device="cuda"), torch.zeros(N, C + 64, device="cuda"), 0, torch.zeros(N, hw, hw, device="cuda"))
            for _ in range(ncons)]
    def multi(): K.gated_stats_multi(x, cg, sg, cons)
    def singles():
        for (ws, m, off, d) in cons:
            K.gated_stats(x, cg, sg, ws, m, off, False, d, True)
    res = []
    for fn in (multi, singles):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 5)
    print(f"{name:6s} K={ncons} C={C} gated={gated}: multi {res[0]:.3f} ms ({x.numel()*2/res[0]/1e6:.0f} GB/s of source)  singles {res[1]:.3f} ms", flush=True)
    del x, cons
