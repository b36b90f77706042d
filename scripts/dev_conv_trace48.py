import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import archs, kernels as K, ttach_compat as tta
cfg = dict(classes=1, decoder_attention_type="scse", decoder_use_batchnorm=True, base_dim=32, encoder_depth=5,
           encoder_name="BoTSER50_Axial_scratch", deep_supervision=False, drop_block_prob=0.0, clf_head=False)
torch.manual_seed(0)
model = archs.get_model("unetplusplusstar", cfg, training=False).to('cuda').eval()
x = torch.randn(6, 3, 1024, 1024, device='cuda')
t = tta.aliases.d4_transform()
os.environ["EDS_CUDA_GRAPHS"] = "0"
for _ in range(2): model.forward_tta(x, t, True)
torch.cuda.synchronize()
K.CONV_TRACE = []
model.forward_tta(x, t, True); torch.cuda.synchronize()
tr, K.CONV_TRACE = K.CONV_TRACE, None
agg = {}
for f, a, b, shp in tr:
    e = agg.setdefault(shp, [0, 0.0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(b); e[2] += f
tot = sum(v[1] for v in agg.values()); fl = sum(v[2] for v in agg.values())
print(f"conv launches {len(tr)} total {tot:.2f} ms, {fl/tot/1e9:.0f} TFLOP/s")
for shp, (n, ms, f) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  N{shp[0]} {shp[1]}x{shp[2]} C{shp[3]}->{shp[4]} k{shp[5]} s{shp[6]}: x{n} {ms:.3f} ms {f/ms/1e9:.0f} TFLOP/s")
