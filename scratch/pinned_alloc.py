"""Dev tool: cost of a fresh page-locked block per call from torch's caching host allocator."""
import time, torch
torch.cuda.init()
n = 13_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for pattern in ("drop", "hold"):
    keep = None
    ts = []
    for i in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
        t1 = time.perf_counter()
        h.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        t2 = time.perf_counter()
        a = h.numpy()
        if pattern == "hold":
            keep = a
        del h, a
        ts.append((round(1e3 * (t1 - t0), 3), round(1e3 * (t2 - t1), 3)))
    print(pattern, ts)
