"""HBM bandwidth by access mix (dev tool): pure read, pure write, copy, 1:4 read:write -- the roofs that bound the
write-heavy kernels (x2 upsampling, 1x1 expansions, paste)."""
import torch
dev = "cuda"
n = 1 << 30          # 1 Gi elements
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
x = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
y = torch.empty(n, dtype=torch.bfloat16, device=dev)
ms = timeit(lambda: y.fill_(1.0)); print(f"pure write  (fill_ 2 GiB)      : {2*n/ms/1e6:7.0f} GB/s")
ms = timeit(lambda: y.zero_()); print(f"pure write  (memset 2 GiB)     : {2*n/ms/1e6:7.0f} GB/s")
ms = timeit(lambda: torch.sum(x)); print(f"pure read   (sum 2 GiB)        : {2*n/ms/1e6:7.0f} GB/s")
ms = timeit(lambda: y.copy_(x)); print(f"copy        (2 GiB -> 2 GiB)   : {4*n/ms/1e6:7.0f} GB/s")
q = x[: n // 4].view(-1, 1)
y4 = y.view(-1, 4)
ms = timeit(lambda: y4.copy_(q.expand(-1, 4))); print(f"1 read : 4 write (expand copy) : {(2*n//4 + 2*n)/ms/1e6:7.0f} GB/s")
