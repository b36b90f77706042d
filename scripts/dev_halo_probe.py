import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
shapes = [(8, 512, 512, 448, 64), (8, 512, 512, 64, 64), (8, 256, 256, 256, 128)]
for (N, H, W, C, Cout) in shapes:
    x = torch.randn(N, H, W, C, device='cuda').bfloat16()
    w = (torch.randn(Cout, 3, 3, C, device='cuda') / math.sqrt(9*C)).bfloat16()
    b = torch.zeros(Cout, device='cuda')
    y = torch.empty(N, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    fl = 2.0 * N * H * W * C * Cout * 9
    ms1 = timeit(lambda: K.conv2d(x, w, b, 1, 1, True, None, out=y, impl='halo'))
    tiles = N * (H // 32) * (W // 8)
    mmas = 3 * (C // 64) * 24
    cyc = ms1 * 1e-3 * 1.9e9 / math.ceil(tiles / 148) / mmas
    print(f"dbg={os.environ.get('EDS_HALO_DEBUG','0')} st={os.environ.get('EDS_HALO_STAGES','-')} C{C}->{Cout} {H}: {ms1:.3f} ms {fl/ms1/1e9:.0f} TF  ~{cyc:.0f} cyc/MMA@1.9GHz", flush=True)
