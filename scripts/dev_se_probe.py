"""SE bottleneck tail: conv3 with scale + residual + ReLU in its epilogue vs conv3 + channel_mean + se_scale_add_relu
(dev tool), 48 maps."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K
N = 48
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, H, Cin, Cout in [("layer1", 256, 64, 256), ("layer2", 128, 128, 512), ("layer3", 64, 256, 1024)]:
    x = torch.randn(N, H, H, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, 1, 1, Cin, device="cuda") / math.sqrt(Cin)).bfloat16()
    b = torch.zeros(Cout, device="cuda")
    res = torch.randn(N, H, H, Cout, device="cuda").bfloat16()
    gate = torch.rand(N, Cout, device="cuda")
    y = torch.empty(N, H, H, Cout, device="cuda", dtype=torch.bfloat16)
    t_plain = timeit(lambda: K.conv2d(x, w, b, 1, 0, False, None, out=y, impl="tc"))
    t_res = timeit(lambda: K.conv2d(x, w, b, 1, 0, True, res, out=y, impl="tc"))
    t_fused = timeit(lambda: K.conv1x1_se(x, w, b, gate, res, out=y))
    t_mean = timeit(lambda: K.channel_mean(y))
    t_scale = timeit(lambda: K.se_scale_add_relu(y, gate, res, out=y))
    t_mean_in = timeit(lambda: K.channel_mean(x))
    gb = (x.numel() + 2 * y.numel()) * 2 / 1e6
    print(f"{name}: plain {t_plain:.3f}  +residual {t_res:.3f}  fused(gate+res) {t_fused:.3f} ms ({gb/t_fused:.0f} GB/s) | "
          f"unfused tail: mean {t_mean:.3f} + scale {t_scale:.3f}; mean(in) {t_mean_in:.3f}", flush=True)
