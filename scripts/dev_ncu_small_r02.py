"""ncu probe, round 2: the small kernels that sit off every roofline at the bench shapes
(8 maps = one 1024^2 tile x 8 D4 views): stem, axial attention, 16-channel tail, head, SE tail on a small map, maxpool."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K, _lib, ttach_compat as tta
dev = "cuda"
torch.manual_seed(0)
S = 1024
aug, _ = tta.view_maps(tta.aliases.d4_transform(), S, S)
x = torch.randn(1, 3, S, S, device=dev)
wp = K.stem_pack_weights(torch.randn(7, 7, 3, 64, device=dev) / 12)
bias = torch.zeros(64, device=dev)
for _ in range(2):
    f0 = K.stem_conv_mma(x, aug, wp, bias)                     # [8, 512, 512, 64]
for _ in range(2):
    K.maxpool2d(f0, 3, 2, 1)
# axial attention at stride 16 (L = 64), 8 heads x 64 value channels, both axes
N, L, heads, dqk, dv = 8, 64, 8, 8, 64
qk = (torch.randn(N, L, L, heads * (2 * dqk + dv), device=dev) * 0.5).bfloat16()
rel = torch.randn(2 * dqk + dv, 2 * L - 1, device=dev) * 0.5
ss = torch.randn(heads, 3, device=dev) * 0.3
osc, osh = torch.randn(2, heads * dv, device=dev), torch.randn(2, heads * dv, device=dev)
for axis in (0, 1):
    K.axial_attention(qk, None, axis, heads, dqk, dv, rel, ss, osc, osh)
# decoder tail at 1024^2: 16 -> 16 (mma.sync) and the 16 -> 1 head
t = torch.randn(8, S, S, 16, device=dev).bfloat16()
w16 = (torch.randn(16, 3, 3, 16, device=dev) / 12).bfloat16()
for _ in range(2):
    K.conv3x3_small(t, w16, torch.zeros(16, device=dev))
for _ in range(2):
    K.head_conv3x3(t, torch.randn(1, 3, 3, 16, device=dev) / 12, torch.zeros(1, device=dev))
# SE tail on a 128^2 map (unfused form): y = relu(x * gate + residual)
a = torch.randn(8, 128, 128, 512, device=dev).bfloat16()
r = torch.randn_like(a)
g = torch.rand(8, 512, device=dev)
for _ in range(2):
    K.se_scale_add_relu(a, g, r)
torch.cuda.synchronize()
print("ok")
