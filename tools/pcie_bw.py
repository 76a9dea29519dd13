#!/usr/bin/env python
"""Pinned host<->device copy bandwidth on this box: the floor of the end-to-end bench numbers.

  python tools/pcie_bw.py            one GPU, both directions
  python tools/pcie_bw.py --all      device->host with 1, 2, 4, 8 ... GPUs copying AT THE SAME TIME into the one host
                                     (what every rank of a multi-GPU end-to-end step does); one JSON line per count
"""
import json
import sys
import time

import torch

n = 1 << 30


def one_gpu():
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): fn()
        e1.record(); torch.cuda.synchronize()
        print("%s %.1f GB/s" % (name, 4 * n / e0.elapsed_time(e1) / 1e6))


def concurrent():
    ng = torch.cuda.device_count()
    hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(ng)]
    ds, ss = [], []
    for g in range(ng):
        with torch.cuda.device(g):
            ds.append(torch.empty(n, dtype=torch.uint8, device="cuda:%d" % g))
            ss.append(torch.cuda.Stream(device=g))
    k = 1
    while k <= ng:
        for direction in ("d2h", "h2d"):
            def go(reps):
                for g in range(k):
                    with torch.cuda.device(g), torch.cuda.stream(ss[g]):
                        for _ in range(reps):
                            (hs[g].copy_(ds[g], non_blocking=True) if direction == "d2h" else ds[g].copy_(hs[g], non_blocking=True))
                for g in range(k):
                    ss[g].synchronize()
            go(1)
            t0 = time.perf_counter()
            go(4)
            sec = time.perf_counter() - t0
            print(json.dumps({"gpus_copying": k, "direction": direction, "aggregate_GBps": 4 * k * n / sec / 1e9,
                              "per_gpu_GBps": 4 * n / sec / 1e9, "bytes_per_gpu": 4 * n}))
        k *= 2


if __name__ == "__main__":
    concurrent() if "--all" in sys.argv else one_gpu()
