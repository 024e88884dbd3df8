#!/usr/bin/env python
"""PCIe bound of the e2e step: 1.93 GB pinned -> device, 1.43 GB device -> pinned, alone and concurrently."""
import time, torch
dev = torch.device("cuda", 0)
up = torch.empty(1_928_966_400, dtype=torch.uint8).pin_memory()
dn = torch.empty(1_434_827_135, dtype=torch.uint8).pin_memory()
d_up = torch.empty_like(up, device=dev); d_dn = torch.empty(dn.numel(), dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d_up.copy_(up, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): dn.copy_(d_dn, non_blocking=True)
def both():
    h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D alone {a:.1f} ms ({up.numel()/a/1e6:.1f} GB/s)  D2H alone {b:.1f} ms ({dn.numel()/b/1e6:.1f} GB/s)  both {c:.1f} ms")
