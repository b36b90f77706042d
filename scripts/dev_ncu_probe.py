"""ncu probe: one launch each of the kernels under study (keep short: ncu replays ~40x)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
torch.manual_seed(0)
if which in ("all", "concat"):
    x = torch.randn(8, 128, 128, 512, device=dev).bfloat16()
    s = torch.randn(8, 256, 256, 512, device=dev).bfloat16()
    w = torch.randn(1024, device=dev)
    for _ in range(2):
        cat, mean, logit = K.concat_stats(x, [s], _lib.UP_BILINEAR, w, 0.1)
    g = torch.rand(8, 1024, device=dev)
    for _ in range(2):
        K.scse_scale(cat, g, logit, out=cat)
    del x, s, cat
if which in ("all", "conv"):
    x = torch.randn(8, 256, 256, 1024, device=dev).bfloat16()
    w = (torch.randn(256, 3, 3, 1024, device=dev) / 96).bfloat16()
    b = torch.zeros(256, device=dev)
    for _ in range(2):
        y = K.conv2d(x, w, b, 1, 1, True, None, impl="tc")
    x = torch.randn(8, 512, 512, 64, device=dev).bfloat16()
    w = (torch.randn(64, 3, 3, 64, device=dev) / 24).bfloat16()
    b = torch.zeros(64, device=dev)
    for _ in range(2):
        y = K.conv2d(x, w, b, 1, 1, True, None, impl="tc")
if which in ("all", "hist"):
    prob = torch.rand(8, 2848 * 4288, device=dev)
    gt = (torch.rand(8, 2848 * 4288, device=dev) < 0.01).to(torch.uint8)
    for _ in range(2):
        h, st = K.pr_hist(prob, gt)
torch.cuda.synchronize()
print("probe ok")
