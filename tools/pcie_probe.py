#!/usr/bin/env python
"""Measures pinned H2D / D2H / bidirectional copy bandwidth on GPU 0: the ceiling of bench.py's e2e number."""
import json
import time

import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def h2d_chunks():
    with torch.cuda.stream(s1):
        for o in range(0, n, 8 << 20):
            d_a[o:o + (8 << 20)].copy_(h_in[o:o + (8 << 20)], non_blocking=True)


res = {"bytes": n, "h2d_gbs": n / timed(h2d) / 1e9, "d2h_gbs": n / timed(d2h) / 1e9, "bidir_each_gbs": n / timed(both) / 1e9,
       "h2d_8MiB_pieces_gbs": n / timed(h2d_chunks) / 1e9}
print(json.dumps(res))
