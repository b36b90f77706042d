"""ncu probe, round 1 final kernels: one launch each at the bench shapes (N = 8 maps unless noted)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K, _lib, ttach_compat as tta
dev = "cuda"
torch.manual_seed(0)
def conv(N, H, C, Cout, R, impl="tc"):
    x = torch.randn(N, H, H, C, device=dev).bfloat16()
    w = (torch.randn(Cout, R, R, C, device=dev) / math.sqrt(R * R * C)).bfloat16()
    b = torch.zeros(Cout, device=dev)
    for _ in range(2):
        K.conv2d(x, w, b, 1, R // 2, True, None, impl=impl)
conv(8, 256, 1024, 256, 3)          # wide decoder layer (generic kernel)
conv(8, 512, 448, 64, 3)            # narrow decoder layer (halo kernel)
conv(8, 512, 64, 64, 3)             # thin layer (halo kernel)
conv(8, 256, 64, 256, 1)            # bottleneck expansion 1x1 (memory bound)
def gsrc(N, h, c, gated=True):
    x = torch.randn(N, h, h, c, device=dev).bfloat16()
    return (x, torch.rand(N, c, device=dev), torch.rand(N, h, h, device=dev)) if gated else (x, None, None)
srcs = [gsrc(8, 256, 256), gsrc(8, 512, 64), gsrc(8, 512, 64), gsrc(8, 512, 64, False)]     # x_1_3
ct = sum(t[0].shape[3] for t in srcs)
wv = torch.randn(ct, device=dev); mean = torch.empty(8, ct, device=dev)
d0 = torch.empty(8, 256, 256, device=dev); d1 = torch.empty(8, 512, 512, device=dev)
off = 0
for k, (t, cg, sg) in enumerate(srcs):
    c = t.shape[3]
    K.gated_stats(t, cg, sg, wv[off:off + c], mean, off, k == 0, d0 if k == 0 else d1, k > 1)
    off += c
sg1 = K.sse_finalize(d0, d1, _lib.UP_BILINEAR, 0.1)
for _ in range(2):
    K.concat_gated(srcs, _lib.UP_BILINEAR, torch.rand(8, ct, device=dev), sg1)
del srcs
# axial attention, encoder block 0 (L = 64, 8 heads, dv 64)
qkv = torch.randn(8, 64, 64, 640, device=dev).bfloat16()
rel = torch.randn(80, 127, device=dev)
for ax in (0, 1):
    K.axial_attention(qkv, None, ax, 8, 8, 64, rel, torch.rand(8, 3, device=dev), torch.rand(2, 512, device=dev), torch.rand(2, 512, device=dev))
# stem (8 views of one tile), merge, paste, preprocess
x = torch.randn(1, 3, 1024, 1024, device=dev)
aug, deaug = tta.view_maps(tta.aliases.d4_transform(), 1024, 1024)
wst = torch.randn(7, 7, 3, 64, device=dev)
K.stem_conv_mma(x, aug, K.stem_pack_weights(wst), torch.zeros(64, device=dev))
logits = torch.randn(8, 6, 1024, 1024, device=dev)
prob = K.tta_merge(logits, deaug, True)
full = torch.zeros(2848, 4288, device=dev)
K.resize_paste(prob[0], full, (0, 0, 1024, 1024), (0, 0), (2048, 2048))
img = torch.randint(0, 256, (2848, 4288, 3), device=dev, dtype=torch.uint8)
K.preprocess_tile(img, 0, 0, 1024, [0.45, 0.22, 0.06], [0.33, 0.17, 0.09])
# scoring: 8 images in one launch
p8 = torch.rand(8, 2848 * 4288, device=dev)
g8 = (torch.rand(8, 2848 * 4288, device=dev) < 0.01).to(torch.uint8)
for _ in range(2):
    h, st = K.pr_hist(p8, g8)
K.pr_scan(h, st)
torch.cuda.synchronize()
print("probe ok")
