#!/usr/bin/env python
"""Timeline probe of the host-buffer path (sky_submit / sky_wait): per-call wall times and kernel_ms."""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

from skyplane_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sub", type=int, default=256)
ap.add_argument("--slots", type=int, default=4)
ap.add_argument("--batches", type=int, default=12)
ap.add_argument("--chunk-mib", type=int, default=8)
a = ap.parse_args()
cb = a.chunk_mib << 20
stride_in = native.round16(cb)
bound = native.frame_bound(cb)
stride_out = native.round16(bound)
ctx = native.Context(0, a.sub * stride_in, a.sub, a.slots)
pool_n = 64
pin_in = native.PinnedBuffer(pool_n * stride_in)
rng = np.random.default_rng(0)
for i in range(pool_n):
    pin_in.view[i * stride_in : i * stride_in + cb] = rng.bytes(cb)
pin_out = [native.PinnedBuffer(a.sub * stride_out) for _ in range(a.slots)]
T0 = time.perf_counter()
ev = []
inflight = []


def pop():
    t, b = inflight.pop(0)
    t0 = time.perf_counter()
    ol, dg, kms = ctx.wait(t)
    ev.append(("wait", b, round((t0 - T0) * 1e3, 1), round((time.perf_counter() - t0) * 1e3, 1), round(kms, 1)))


for b in range(a.batches):
    if len(inflight) == a.slots:
        pop()
    src = [pin_in.addr + ((i % pool_n) * stride_in) for i in range(a.sub)]
    dst = [pin_out[b % a.slots].addr + i * stride_out for i in range(a.sub)]
    t0 = time.perf_counter()
    t = ctx.submit(src, [cb] * a.sub, dst, [bound] * a.sub)
    ev.append(("submit", b, round((t0 - T0) * 1e3, 1), round((time.perf_counter() - t0) * 1e3, 1), None))
    inflight.append((t, b))
while inflight:
    pop()
total = time.perf_counter() - T0
print(json.dumps({"sub": a.sub, "slots": a.slots, "batches": a.batches, "gbs": a.batches * a.sub * cb / total / 1e9, "total_ms": total * 1e3}))
for e in ev:
    print(e)
