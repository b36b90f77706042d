"""Time the upsampled part of the decoder concat (PART 0) on the benchmark's layer shapes (dev tool).
Run twice: EDS_UPSAMPLE_TILED=0 / 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K, _lib
N = 48
shapes = [("x_1_3 up", 256, 256, 256, 448, True), ("x_0_3 up", 256, 256, 64, 320, True), ("x_3_3 up", 256, 256, 256, 320, False),
          ("x_0_4 up", 512, 512, 32, 32, True), ("x_1_2 up", 128, 128, 512, 1024, False), ("x_0_2 up", 128, 128, 128, 896, False)]
for name, h, w, C0, Ctot, gated in shapes:
    x = torch.randn(N, h, w, C0, device="cuda").bfloat16()
    cg0 = torch.rand(N, C0, device="cuda") if gated else None
    sg0 = torch.rand(N, h, w, device="cuda") if gated else None
    cg1 = torch.rand(N, Ctot, device="cuda")
    sg1 = torch.rand(N, 2 * h, 2 * w, device="cuda")
    # split destination: stride C0 (dense upsampled map), gate vector of the whole concat
    skip = torch.zeros(N, 2 * h, 2 * w, 16, device="cuda").bfloat16()
    out = torch.empty(N, 2 * h, 2 * w, C0, device="cuda", dtype=torch.bfloat16)
    fn = lambda: K.concat_gated([(x, cg0, sg0)], _lib.UP_BILINEAR, cg1[:, :C0].contiguous(), sg1, out=out)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    byts = x.numel() * 2 + out.numel() * 2
    print(f"tiled={os.environ.get('EDS_UPSAMPLE_TILED','1')} {name:9s} {h}x{w} C0={C0} gated={gated}: {ms:.3f} ms  {byts/ms/1e6:.0f} GB/s", flush=True)
    del x, out, sg1, skip
