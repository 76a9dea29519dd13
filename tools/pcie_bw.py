#!/usr/bin/env python
"""Measures pinned host<->device copy bandwidth on this box (the floor for the end-to-end bench numbers)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): fn()
    e1.record(); torch.cuda.synchronize()
    print("%s %.1f GB/s" % (name, 4 * n / e0.elapsed_time(e1) / 1e6))
