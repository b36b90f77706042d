"""Per-rank step timing under torchrun (diagnostic): GPU time and host enqueue time of bench.step_resident."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from eyediseasesegmentation_b200 import archs, kernels as K, _driver as drv, ttach_compat as tta
from eyediseasesegmentation_b200.archs import get_preprocessing_fn
rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
use_nccl = os.environ.get("PROBE_NCCL", "1") == "1"
if world > 1 and use_nccl:
    dist.init_process_group("nccl", device_id=dev)
_, mean, std = get_preprocessing_fn("IDRiD", False)
tfm = tta.aliases.d4_transform()
models = []
for i in range(4):
    torch.manual_seed(1999 + i)
    m = archs.get_model("unetplusplusstar", bench.star_cfg(32), training=False).to(dev).eval(); m.precision = "bf16"; models.append(m)
img, masks = bench.synth_image(rank)
image = torch.from_numpy(img).to(dev); gts = [torch.from_numpy(masks[l]).to(dev) for l in bench.LESIONS]
def step():
    for m, gt in zip(models, gts):
        preds = drv.tiled_probability_map(m, tfm, image, 1024, mean, std)
        h, s = K.pr_hist(preds.view(1, -1), gt.view(1, -1)); K.pr_scan(h, s)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"rank {rank} dev {torch.cuda.current_device()} step {i}: gpu {a.elapsed_time(b):.1f} ms host {1e3*(t1-t0):.1f} ms cpus {len(os.sched_getaffinity(0))} nccl {use_nccl}", flush=True)
