"""tta_merge under different view sets + a same-size copy, L2 flushed between repetitions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eyediseasesegmentation_b200 import kernels as K, ttach_compat as tta

dev = torch.device("cuda", 0)
V, B, S = 8, 6, 1024
_, deaug = tta.view_maps(tta.aliases.d4_transform(), S, S)
print("deaug maps", deaug)
logits = torch.randn((V, B, S, S), device=dev)
prob = torch.empty((B, S, S), device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ident = [deaug[0]] * 8
transp = [m for m in deaug if m[1] != 0]
transp = (transp * 8)[:8]
nbytes = V * B * S * S * 4 + B * S * S * 4


def timed(fn, label, nb):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[3]
    print(f"{label:28s} {ms*1e3:8.1f} us  {nb/ms/1e6:8.1f} GB/s")


timed(lambda: K.tta_merge(logits, deaug, True, out=prob), "merge d4", nbytes)
timed(lambda: K.tta_merge(logits, ident, True, out=prob), "merge identity x8", nbytes)
timed(lambda: K.tta_merge(logits, transp, True, out=prob), "merge transposed x8", nbytes)
src = logits.view(-1)[: nbytes // 8]
dst = torch.empty_like(src)
timed(lambda: dst.copy_(src), "copy same bytes", nbytes)
timed(lambda: torch.sum(logits, dim=0, out=prob), "torch.sum over views", nbytes)
big = torch.empty(1 << 28, device=dev); big2 = torch.empty_like(big)
timed(lambda: big2.copy_(big), "copy 2 GiB", 2 * big.numel() * 4)
