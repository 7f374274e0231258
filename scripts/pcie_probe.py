#!/usr/bin/env python
"""PCIe probe for the host round trip: H2D alone, D2H alone, both at once on two streams (25 MB each way)."""
import torch
n = 25165824
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
import time
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 20 * 1e3
    print("%s: %.3f ms per 25 MB (each way)  -> %.1f GB/s per direction" % (name, ms, n / ms / 1e6))
