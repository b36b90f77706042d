"""Kernel-time breakdown of one proposed-net forward at the bench batch (6 tiles x 8 views = 48 maps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import archs, kernels as K, ttach_compat as tta
B = int(os.environ.get("BATCH", "6"))
cfg = dict(classes=1, decoder_attention_type="scse", decoder_use_batchnorm=True, base_dim=32, encoder_depth=5,
           encoder_name="BoTSER50_Axial_scratch", deep_supervision=False, drop_block_prob=0.0, clf_head=False)
torch.manual_seed(0)
model = archs.get_model("unetplusplusstar", cfg, training=False).to('cuda').eval()
x = torch.randn(B, 3, 1024, 1024, device='cuda')
t = tta.aliases.d4_transform()
for _ in range(2):
    model.forward_tta(x, t, True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); model.forward_tta(x, t, True); b.record(); torch.cuda.synchronize()
print(f"forward_tta batch {B} x 8 views: {a.elapsed_time(b):.2f} ms -> {B*8*1872.19/a.elapsed_time(b):.1f} TFLOP/s algorithmic")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.forward_tta(x, t, True); torch.cuda.synchronize()
rows = [(e.key, e.self_device_time_total / 1e3, e.count) for e in prof.key_averages()]
tot = sum(r[1] for r in rows)
for k, ms, n in sorted(rows, key=lambda r: -r[1])[:30]:
    print(f"{ms:9.3f} ms {100*ms/tot:5.1f}% x{n:<3d} {k[:90]}")
print(f"total {tot:.2f} ms")
