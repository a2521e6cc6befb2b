import torch, time
n = 2 * 1024**3
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for _ in range(2):
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print('D2H pinned GB/s', 3 * n / (time.perf_counter() - t0) / 1e9)
t0 = time.perf_counter()
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
print('H2D pinned GB/s', 3 * n / (time.perf_counter() - t0) / 1e9)
# numpy copy speed out of pinned memory (what a consumer of the records pays)
import numpy as np
a = h.numpy()
t0 = time.perf_counter(); b = a.copy(); print('host memcpy GB/s', n / (time.perf_counter() - t0) / 1e9)
