"""pinned host -> device copy bandwidth of this box (what bounds bench.py's e2e leg)"""
import torch, time
n = 1 << 31
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D pinned 2 GiB: {n / dt / 1e9:.1f} GB/s")
h2 = h[: 4096 * 131072 * 2].view(4096, -1)
t0 = time.perf_counter()
for _ in range(5):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"D2H pinned 2 GiB: {n / dt / 1e9:.1f} GB/s")
