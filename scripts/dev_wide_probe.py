"""Time the narrow-output 3x3 layers of the proposed net (48 maps) on the halo and the dw-grouped wide kernel."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K

dev = "cuda"
N = 48
LAYERS = [  # H, Cin, Cout
    (256, 896, 64), (256, 64, 64), (512, 320, 32), (512, 32, 32), (512, 448, 64), (512, 384, 64), (512, 320, 64),
    (512, 64, 64), (1024, 32, 16), (1024, 16, 16), (512, 128, 32),
]
impls = sys.argv[1:] or ["halo", "wide", "tc"]
for H, C, Co in LAYERS:
    n = N if H < 1024 else 24
    x = torch.randn(n, H, H, C, device=dev).bfloat16()
    w = (torch.randn(Co, 3, 3, C, device=dev) / math.sqrt(9 * C)).bfloat16()
    b = torch.randn(Co, device=dev)
    y = torch.empty(n, H, H, Co, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * n * H * H * Co * 9 * C
    line = f"{H:5d} {C:4d}->{Co:3d}"
    for impl in impls:
        for _ in range(0 if os.environ.get("PROBE_ONE_LAUNCH") else 2):     # PROBE_ONE_LAUNCH=1: one launch per layer (ncu)
            K.conv2d(x, w, b, 1, 1, True, None, out=y, impl=impl)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 1 if os.environ.get("PROBE_ONE_LAUNCH") else 3
        for _ in range(reps):
            K.conv2d(x, w, b, 1, 1, True, None, out=y, impl=impl)
        e.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(e) / reps
        line += f" | {impl} {ms:7.3f} ms {flops / ms / 1e9:7.0f} TF/s"
    print(line, flush=True)
    del x, y
