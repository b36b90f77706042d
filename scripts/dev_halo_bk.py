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
N = 48
for (H, C, Cout) in [(512, 64, 64), (256, 64, 64), (512, 448, 64), (512, 320, 32), (128, 128, 128)]:
    x = torch.randn(N, H, H, C, device='cuda').bfloat16()
    w = (torch.randn(Cout, 3, 3, C, device='cuda') / math.sqrt(9*C)).bfloat16()
    b = torch.zeros(Cout, device='cuda')
    y = torch.empty(N, H, H, Cout, device='cuda', dtype=torch.bfloat16)
    ms = timeit(lambda: K.conv2d(x, w, b, 1, 1, True, None, out=y, impl='halo'))
    fl = 2.0 * N * H * H * C * Cout * 9
    print(f"bk={os.environ.get('EDS_HALO_BK','64')} st={os.environ.get('EDS_HALO_STAGES','-')} N{N} {H} C{C}->{Cout}: {ms:.3f} ms {fl/ms/1e9:.0f} TF", flush=True)
    del x, w, y
