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
import time
for i in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); model.forward_tta(x, t, True); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"graphs={os.environ.get('EDS_CUDA_GRAPHS','1')} call {i}: gpu {a.elapsed_time(b):.2f} ms, host enqueue {1e3*(t1-t0):.2f} ms", flush=True)
print("max mem GB", torch.cuda.max_memory_allocated()/2**30, "reserved", torch.cuda.memory_reserved()/2**30)
