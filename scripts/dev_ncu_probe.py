"""ncu probe: one or two launches each of the kernels under study (keep short: ncu replays ~40x).

    python scripts/dev_ncu_probe.py [concat|conv|hist|gated|all]
"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda"
torch.manual_seed(0)


def timed(name, fn, nbytes, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"{name}: {ms*1e3:.1f} us, {nbytes/ms/1e6:.0f} GB/s algorithmic", flush=True)


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
if which in ("all", "gated"):
    def gsrc(N, h, w, c, gated=True):
        x = torch.randn(N, h, w, c, device=dev).bfloat16()
        return (x, torch.rand(N, c, device=dev), torch.rand(N, h, w, device=dev)) if gated else (x, None, None)
    for name, x0, skips in (("x_1_2", gsrc(8, 128, 128, 512, False), [gsrc(8, 256, 256, 512, False)]),
                            ("x_1_3", gsrc(8, 256, 256, 256), [gsrc(8, 512, 512, 64), gsrc(8, 512, 512, 64), gsrc(8, 512, 512, 64, False)])):
        srcs = [x0] + skips
        N, h, w_, _ = x0[0].shape
        ct = sum(t[0].shape[3] for t in srcs)
        wv = torch.randn(ct, device=dev)
        mean = torch.empty(N, ct, device=dev)
        d0 = torch.empty(N, h, w_, device=dev); d1 = torch.empty(N, 2 * h, 2 * w_, device=dev)
        off = 0
        for k, (t, cg, sg) in enumerate(srcs):
            c = t.shape[3]
            timed(f"{name} gated_stats src{k} C{c} {t.shape[1]}^2", lambda: K.gated_stats(t, cg, sg, wv[off:off + c], mean, off, k == 0, d0 if k == 0 else d1, k > 1), t.numel() * 2)
            off += c
        cg1 = torch.rand(N, ct, device=dev)
        sg1 = K.sse_finalize(d0, d1, _lib.UP_BILINEAR, 0.1)
        out = torch.empty(N, 2 * h, 2 * w_, ct, device=dev, dtype=torch.bfloat16)
        timed(f"{name} concat_gated Ct{ct}", lambda: K.concat_gated(srcs, _lib.UP_BILINEAR, cg1, sg1, out=out),
              sum(t[0].numel() for t in srcs) * 2 + out.numel() * 2)
        del srcs, out, x0, skips
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
