"""Plain pinned-memory copy bandwidth of the box (the ceiling of bench.py's e2e): python tools/pcie_probe.py"""
import time
import torch

n = 2 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for name, both in (("d2h alone", False), ("d2h with h2d traffic", True)):
    for _ in range(2):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(name, round(4 * n / dt / 1e9, 2), "GB/s")
# the e2e shape: 512 MB pieces from two device slots into a 4 GiB pinned range
big = torch.empty(4 << 30, dtype=torch.uint8).pin_memory()
slots = [torch.empty(512 << 20, dtype=torch.uint8, device="cuda") for _ in range(2)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(8):
        big[k * (512 << 20):(k + 1) * (512 << 20)].copy_(slots[k & 1], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("8 x 512 MB into a 4 GiB pinned range", round((4 << 30) / dt / 1e9, 2), "GB/s")
