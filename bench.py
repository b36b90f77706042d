#!/usr/bin/env python
"""Benchmark of the inference-and-scoring hot path (BASELINE.json metric).

Workload ("cfg4"): synthetic IDRiD-shaped 2848x4288 RGB fundus images, four
single-class lesion models (EX/HE/MA/SE) of the proposed UNet++* (base_dim 32), sliding window
(2048 px windows -> 1024^2 tiles, 6 tiles per image), D4 8-view TTA, sigmoid, x2 bilinear paste,
per-image PR/ROC histogram + scan.  One step = one image through all four lesion models
(6 * 8 * 4 = 192 network forwards = 359.5 algorithmic TFLOP).  metric = images per second.

  python bench.py [--gpus N --steps K --warmup W]          B200 path (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                   the reference's CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

H, W, S = 2848, 4288, 1024
LESIONS = ("EX", "HE", "MA", "SE")
PREVALENCE = {"EX": 0.01, "HE": 0.01, "MA": 0.001, "SE": 0.005}
TILES, VIEWS = 6, 8
GFLOP_PER_FORWARD = 1872.19                      # BASELINE.md section 2 (reference modules, FLOP = 2*MAC)
TFLOP_PER_IMAGE = TILES * VIEWS * len(LESIONS) * GFLOP_PER_FORWARD / 1e3
METRIC = "IDRiD-size (2848x4288) fundus images/s, 4 lesion models, sliding-window + 8-view D4 TTA + AUC-PR"


def star_cfg(base_dim=32):
    return dict(classes=1, decoder_attention_type="scse", decoder_use_batchnorm=True, base_dim=base_dim,
                encoder_depth=5, encoder_name="BoTSER50_Axial_scratch", deep_supervision=False,
                drop_block_prob=0.0, clf_head=False)


def synth_image(seed):
    """uint8 noise inside a centred disc of radius 1400, black outside (fundus-like), + 4 blob masks."""
    import numpy as np
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    yy, xx = np.ogrid[:H, :W]
    disc = (yy - H / 2) ** 2 + (xx - W / 2) ** 2 <= 1400 ** 2
    img[~disc] = 0
    masks = {}
    for i, les in enumerate(LESIONS):
        coarse = rng.random((H // 16, W // 16)) < PREVALENCE[les]           # Bernoulli blobs of 16x16 px
        m = np.kron(coarse, np.ones((16, 16), dtype=bool)) & disc
        masks[les] = m.astype(np.uint8)
    return img, masks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ B200 arm
def _claim_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line there.
    Point fd 1 at stderr for the whole run and return a handle on the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def run_b200(args):
    real_stdout = _claim_stdout()
    import numpy as np
    import torch
    import torch.distributed as dist
    from eyediseasesegmentation_b200 import archs, kernels as K, _driver as drv, ttach_compat as tta
    from eyediseasesegmentation_b200.aucpr import score_device
    from eyediseasesegmentation_b200.archs import get_preprocessing_fn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    _, mean, std = get_preprocessing_fn("IDRiD", False)
    tfm = tta.aliases.d4_transform()
    models = {}
    for i, les in enumerate(LESIONS):
        torch.manual_seed(1999 + i)
        m = archs.get_model("unetplusplusstar", star_cfg(32), training=False).to(dev).eval()
        m.precision = "bf16"
        m.engine()
        models[les] = m

    # this rank's images (weak scaling: every rank runs `steps` images of its own)
    n_img = args.warmup + args.steps
    distinct = min(n_img, 2)                      # two distinct synthetic images, reused round-robin
    host_imgs, host_masks = [], []
    for j in range(distinct):
        img, masks = synth_image(1000 * rank + j)
        host_imgs.append(torch.from_numpy(img).pin_memory())
        host_masks.append({les: torch.from_numpy(masks[les]).pin_memory() for les in LESIONS})
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_masks = [{les: t.to(dev) for les, t in d.items()} for d in host_masks]

    def step_resident(j):
        """one image, inputs already in HBM, results stay on the device (value)."""
        out = []
        for les in LESIONS:
            preds = drv.tiled_probability_map(models[les], tfm, dev_imgs[j], S, mean, std, tiles_per_batch=args.tiles)
            hist, strad = K.pr_hist(preds.view(1, -1), dev_masks[j][les].view(1, -1))
            out.append(K.pr_scan(hist, strad))
        return out

    def step_e2e(j):
        """same through the host-facing unit of tta_patches (`_driver.infer_image_host`): pinned host
        image + mask -> device, probability map and scores back into pinned host memory, one stream
        synchronisation per lesion map (what tta_patches' generator yields per image)."""
        res = []
        for les in LESIONS:
            pred, scores = drv.infer_image_host(models[les], tfm, host_imgs[j], host_masks[j][les], S, mean, std,
                                                tiles_per_batch=args.tiles, copy=False, slot=les)
            res.append((pred, scores.ap))
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i % distinct)
        if world > 1:                              # warm the exact collective of the timed region (lazy
            warm = torch.zeros(2 * 19 + 2, dtype=torch.int64, device=dev)   # NCCL kernel load / channel set-up)
            dist.all_reduce(warm)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        pooled = torch.zeros(2 * 19 + 2, dtype=torch.int64, device=dev)
        for i in range(steps):
            out = fn((warmup + i) % distinct)
            if world > 1 and out and isinstance(out[0], tuple) and len(out[0]) == 4:
                for (_ap, _roc, counts, totals) in out:      # this rank's pooled (tp, pp) per threshold + totals
                    pooled[:38] += counts.reshape(-1)
                    pooled[38:] += totals.reshape(-1)
        if world > 1:                              # the path's single collective: pooled integer counts
            dist.all_reduce(pooled)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local) if rank == 0 else None
    launch_marks = []

    def counted(j):
        before = K.LAUNCHES[0]
        out = step_resident(j)
        launch_marks.append(K.LAUNCHES[0] - before)
        return out

    ms_total = timed(counted, args.steps, args.warmup)
    launches = sum(launch_marks[-args.steps:])             # kernels of libeds_b200 launched in the timed steps
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 3))

    line = None
    if rank == 0:
        value = world * args.steps / (ms_total / 1e3)
        e2e = world * args.steps / (ms_e2e / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json, sustained)" if peaks else "fallback (B200_PROFILING.md)"
        roof = conv_roofline(models["EX"], tfm, dev_imgs[0], mean, std, args.tiles, tensor_peak, peak_src)
        hist_roof = hist_roofline(dev, hbm_peak, peak_src)
        blend_roof = blend_roofline(dev, hbm_peak, peak_src)
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg4: 2848x4288 images x 4 lesion models (proposed UNet++*, base_dim 32, random "
                                   "init), 6 tiles of 1024^2 x 8 D4 views each, sigmoid + x2 paste + PR histogram/scan",
                       "tiles_per_batch": args.tiles, "algorithmic_tflop_per_image": TFLOP_PER_IMAGE,
                       "achieved_algorithmic_tflops_per_gpu": TFLOP_PER_IMAGE * value / world,
                       "l2": "working set per step (>5 GB of activations per tile batch) far exceeds the 126 MB L2",
                       "parallelism": f"images sharded over {world} rank(s); one int64 all-reduce of pooled counts"},
            "e2e": {"value": e2e, "unit": "images/s",
                    "h2d_bytes_per_step": len(LESIONS) * (H * W * 3 + H * W),
                    "d2h_bytes_per_step": len(LESIONS) * (H * W * 4 + 19 * 2 * 8 + 2 * 8 + 16)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_hist": hist_roof,
            "roofline_blend": blend_roof,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(sample_only=True)
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def conv_roofline(model, tfm, image, mean, std, tiles, peak, peak_src):
    """Dominant kernel = conv_igemm_kernel.  One instrumented tile batch (same input as the timed
    steps) with CUDA events around every launch of that kernel on the launching stream; achieved =
    algorithmic FLOPs of those launches (2*M*Cout*K with M = real output pixels) / their summed time."""
    import torch
    from eyediseasesegmentation_b200 import kernels as K, _driver as drv
    drv.tiled_probability_map(model, tfm, image, S, mean, std, tiles_per_batch=tiles)   # warm
    torch.cuda.synchronize()
    K.CONV_TRACE = []
    drv.tiled_probability_map(model, tfm, image, S, mean, std, tiles_per_batch=tiles)
    torch.cuda.synchronize()
    trace, K.CONV_TRACE = K.CONV_TRACE, None
    flops = sum(t[0] for t in trace)
    ms = sum(t[1].elapsed_time(t[2]) for t in trace)
    achieved = flops / (ms / 1e3) / 1e12
    return {"kernel": "conv_igemm_kernel + conv3x3_halo_kernel (tcgen05 implicit GEMM, all 98 conv launches of a pass)", "bound": "tensor", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
            "launches": len(trace), "avg_launch_ms": ms / max(1, len(trace)), "peak_source": peak_src,
            "flops_counted": flops}


def hist_roofline(dev, peak, peak_src):
    """AUC-PR kernel GB/s: the whole 27-image test set in one launch, algorithmic bytes =
    n_px * (4 B score + 1 B label)."""
    import torch
    from eyediseasesegmentation_b200 import kernels as K
    n_img = 27
    prob = torch.rand((n_img, H * W), device=dev)
    gt = (torch.rand((n_img, H * W), device=dev) < 0.01).to(torch.uint8)
    hist, strad = K.pr_hist(prob, gt)
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        hist.zero_(); strad.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        K.pr_hist(prob, gt, hist, strad)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = sorted(times)[len(times) // 2]
    gbs = n_img * H * W * 5 / (ms / 1e3) / 1e9
    return {"kernel": "pr_hist_kernel", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
            "frac": gbs / peak, "traffic": None, "launch_ms": ms, "images_per_launch": n_img, "peak_source": peak_src}


def blend_roofline(dev, peak, peak_src):
    """TTA merge (de-augment + mean + sigmoid over V*B logit maps) and the x2 bilinear overwrite-paste of
    the B tiles, at the bench shape (V=8, B=6, S=1024) for the 4 lesion models of one image, back to back as
    in a step (4 x 201 MB of logits: larger than L2, and L2 is flushed before each repetition).
    Algorithmic bytes per launch = V*B*S^2*4 read + B*S^2*4 write for the merge; B*S^2*4 read + H*W*4 write
    (every pixel of the image once) for the paste."""
    import torch
    from eyediseasesegmentation_b200 import kernels as K, ttach_compat as tta
    from eyediseasesegmentation_b200.util import make_grid
    V, B, M = VIEWS, TILES, len(LESIONS)
    _, deaug = tta.view_maps(tta.aliases.d4_transform(), S, S)
    logits = [torch.randn((V, B, S, S), device=dev) for _ in range(M)]
    prob = [torch.empty((B, S, S), device=dev) for _ in range(M)]
    preds = [torch.zeros((H, W), device=dev) for _ in range(M)]
    slices = make_grid((H, W), window=2 * S, min_overlap=32)
    origins = [(int(x1), int(y1)) for (x1, _, y1, _) in slices]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)        # > L2

    for m in range(M):
        K.tta_merge(logits[m], deaug, True, out=prob[m])
        K.paste_tiles_x2(prob[m], preds[m], origins)
    torch.cuda.synchronize()
    t_merge, t_all = [], []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for m in range(M):
            K.tta_merge(logits[m], deaug, True, out=prob[m])
        b.record()
        torch.cuda.synchronize()
        t_merge.append(a.elapsed_time(b) / M)
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for m in range(M):
            K.tta_merge(logits[m], deaug, True, out=prob[m])
            K.paste_tiles_x2(prob[m], preds[m], origins)
        b.record()
        torch.cuda.synchronize()
        t_all.append(a.elapsed_time(b) / M)
    ms_merge, ms_all = sorted(t_merge)[2], sorted(t_all)[2]
    bytes_merge = V * B * S * S * 4 + B * S * S * 4
    bytes_paste = B * S * S * 4 + H * W * 4
    gbs = bytes_merge / (ms_merge / 1e3) / 1e9
    return {"kernel": "tta_merge64_kernel", "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
            "frac": gbs / peak, "traffic": 219.0e6, "traffic_source": "profiles/r01_blend_full.md (dram rd + wr per launch)",
            "launch_ms": ms_merge, "launches_timed": M, "peak_source": peak_src,
            "l2": "256 MB written before each repetition (flush); 4 x 201 MB of logits per repetition",
            "with_paste": {"kernels": "tta_merge64_kernel + paste_tiles_x2_kernel (6 tiles, one launch)", "ms": ms_all,
                           "achieved": (bytes_merge + bytes_paste) / (ms_all / 1e3) / 1e9, "unit": "GB/s"}}


# ------------------------------------------------------------------------------ CPU baseline / reference arm
def cpu_sample_seconds(threads):
    """One bounded sample of the workload on the host cores with the oracle (the reference's PyTorch
    path restated, oracle/nets.py): ONE 1024^2 forward (1 of the 192 per image) + sklearn AP on a
    quarter image."""
    import numpy as np
    import torch
    from oracle import nets, scoring
    import helpers
    torch.set_num_threads(threads)
    model = helpers.build_product_model("unetplusplusstar", star_cfg(32), seed=1999)
    sd = model.state_dict()
    x = torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(0))
    t0 = time.perf_counter()
    with torch.no_grad():
        nets.unetplusplusstar_forward(sd, x, 32)
    t_fwd = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    n = H * W // 4
    prob = rng.random(n, dtype=np.float32)
    gt = (rng.random(n) < 0.01).astype(np.uint8)
    t0 = time.perf_counter()
    scoring.get_auc([(prob, gt, "q")])
    scoring.threshold_counts(prob.reshape(1, -1), gt.reshape(1, -1))
    t_score = (time.perf_counter() - t0) * 4
    return t_fwd, t_score


def cpu_baseline(sample_only=False):
    import torch
    threads = os.cpu_count() or 1
    t_fwd, t_score = cpu_sample_seconds(threads)
    per_image = TILES * VIEWS * len(LESIONS) * t_fwd + len(LESIONS) * t_score
    return {"value": 1.0 / per_image, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"oracle (reference PyTorch path restated) on CPU fp32: one 1024^2 proposed-net forward "
                      f"({t_fwd:.2f} s) extrapolated x192 forwards per image + sklearn AP and 19-threshold counts on a "
                      f"quarter image ({t_score / 4:.2f} s) extrapolated x4 x4 lesions; the reference additionally "
                      f"repeats inference 3x (tta.py:218,221,225), not counted"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (here the oracle port;
    /root/reference does not exist on the GPU box) on all host threads.  Each step is the bounded
    sample of cpu_sample_seconds(); images/s is extrapolated from it (stated in cpu_baseline.sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_sample_seconds(threads)
    per_image = []
    t_start = time.perf_counter()
    for _ in range(args.steps):
        t_fwd, t_score = cpu_sample_seconds(threads)
        per_image.append(TILES * VIEWS * len(LESIONS) * t_fwd + len(LESIONS) * t_score)
        if time.perf_counter() - t_start > 240:
            break
    sec = sum(per_image) / len(per_image)
    value = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": len(per_image), "warmup": min(args.warmup, 1),
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg4 (same as the B200 arm); each step = one 1024^2 forward + quarter-image "
                                   "scoring on the host, extrapolated to a full image (x192 forwards, x16 scoring)"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": "one 1024^2 oracle forward + quarter-image sklearn scoring per step, extrapolated"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles", type=int, default=6, help="tiles per forward batch (x8 views)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
