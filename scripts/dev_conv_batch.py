import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K
def timeit(fn, n=4, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
N = int(os.environ.get("NB", "48"))
for (H, C, Cout, R) in [(512, 448, 64, 3), (512, 320, 32, 3), (512, 64, 64, 3), (256, 1024, 256, 3), (256, 256, 256, 3), (256, 64, 256, 1)]:
    x = torch.randn(N, H, H, C, device='cuda').bfloat16()
    w = (torch.randn(Cout, R, R, C, device='cuda') / math.sqrt(R*R*C)).bfloat16()
    b = torch.zeros(Cout, device='cuda')
    y = torch.empty(N, H, H, Cout, device='cuda', dtype=torch.bfloat16)
    ms = timeit(lambda: K.conv2d(x, w, b, 1, R // 2, True, None, out=y, impl='tc'))
    fl = 2.0 * N * H * H * C * Cout * R * R
    print(f"promo={os.environ.get('EDS_L2_PROMO','256')} N{N} {H} C{C}->{Cout} k{R}: {ms:.3f} ms {fl/ms/1e9:.0f} TF", flush=True)
    del x, w, y
