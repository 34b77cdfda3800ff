"""Measure pinned device->host copy bandwidth (context for bench.py's e2e number)."""
import time
import torch
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(2):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"D2H pinned 1 GiB: {n / dt / 1e9:.1f} GB/s")
t0 = time.perf_counter()
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D pinned 1 GiB: {n / dt / 1e9:.1f} GB/s")
