"""ncu probe, round 2: one or two launches of every hot kernel at the bench shapes (N = 8 maps unless noted)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K, _lib, partition, ttach_compat as tta
from eyediseasesegmentation_b200.util import make_grid
dev = "cuda"
torch.manual_seed(0)
def conv(N, H, C, Cout, R):
    x = torch.randn(N, H, H, C, device=dev).bfloat16()
    w = (torch.randn(Cout, R, R, C, device=dev) / math.sqrt(R * R * C)).bfloat16()
    b = torch.zeros(Cout, device=dev)
    for _ in range(2):
        K.conv2d(x, w, b, 1, R // 2, True, None, impl="tc")
conv(8, 256, 1024, 256, 3)          # conv_igemm: wide decoder layer
conv(8, 512, 448, 64, 3)            # conv3x3_wide: narrow decoder layer
conv(8, 128, 128, 128, 3)           # conv3x3_halo
conv(8, 256, 64, 256, 1)            # 1x1 expansion (write-bound)
conv(8, 256, 256, 64, 1)            # 1x1 reduction (read-bound)
def gsrc(N, h, c, gated=True):
    x = torch.randn(N, h, h, c, device=dev).bfloat16()
    return (x, torch.rand(N, c, device=dev), torch.rand(N, h, h, device=dev)) if gated else (x, None, None)
srcs = [gsrc(8, 256, 256), gsrc(8, 512, 64), gsrc(8, 512, 64), gsrc(8, 512, 64, False)]     # x_1_3
ct = sum(t[0].shape[3] for t in srcs)
for _ in range(2):
    K.concat_gated_split(srcs, _lib.UP_BILINEAR, torch.rand(8, ct, device=dev), torch.rand(8, 512, 512, device=dev))
x, cg, sg = srcs[1]
cons = [(torch.randn(64, device=dev), torch.zeros(8, 320, device=dev), 0, torch.zeros(8, 512, 512, device=dev)) for _ in range(3)]
for _ in range(2):
    K.gated_stats_multi(x, cg, sg, cons)        # x_3_3's output read once for its three consumers
K.gated_stats(x, None, None, torch.randn(64, device=dev), torch.zeros(8, 64, device=dev), 0, False,
              torch.zeros(8, 512, 512, device=dev), False)     # attention2 of a block output
del srcs, cons
# the blend: 6 tiles x 8 views -> preds, one kernel; and its two-kernel form
H, W, S = 2848, 4288, 1024
_, deaug = tta.view_maps(tta.aliases.d4_transform(), S, S)
origins = [(int(a), int(c)) for (a, _, c, _) in make_grid((H, W), window=2 * S, min_overlap=32)]
logits = torch.randn(8, 6, S, S, device=dev)
preds = torch.zeros(H, W, device=dev)
for _ in range(2):
    K.tta_blend_x2(logits, deaug, preds, origins)
prob = K.tta_merge(logits, deaug, True)
K.paste_tiles_x2(prob, preds, origins)
# scoring: 8 images in one launch; one rank's rectangles of one image
p8 = torch.rand(8, H * W, device=dev)
g8 = (torch.rand(8, H * W, device=dev) < 0.01).to(torch.uint8)
for _ in range(2):
    h, st = K.pr_hist(p8, g8)
K.pr_scan(h, st)
cells = partition.owned_cells([(y, y + 2 * S, x, x + 2 * S) for (y, x) in origins], (H, W))
K.pr_hist_rects(p8[0].view(H, W), g8[0].view(H, W), cells[0] + cells[3], h[0], st[0])
torch.cuda.synchronize()
print("probe ok")
