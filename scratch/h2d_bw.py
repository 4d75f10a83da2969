"""Dev tool: pinned host->device copy bandwidth of one 229 MB block (the e2e step's certainty planes + images), alone."""
import torch, time
n = 229_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D {n/1e6:.0f} MB: {ms:.3f} ms = {n/ms/1e6:.1f} GB/s")
h2 = torch.empty(13_000_000, dtype=torch.uint8).pin_memory(); d2 = torch.empty(13_000_000, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
e0.record()
for _ in range(10):
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"H2D with a concurrent 13 MB D2H per copy: {e0.elapsed_time(e1)/10:.3f} ms")
