"""Helpers shared by the GPU parity tests: drive libskychunk through its C ABI."""
from __future__ import annotations

from typing import List, Sequence

from skyplane_b200 import native


def run_device(ctx: native.Context, chunks: Sequence[bytes], flags: int = 0, stream: int = 0):
    """Device-resident path: stage chunks into HBM, call sky_process_device, fetch frames.
    -> (frames: list[bytes], digests: list[bytes], out_lens, kernel_ms)"""
    src_off, dst_off, caps = [], [], []
    ip = op = 0
    for c in chunks:
        src_off.append(ip)
        dst_off.append(op)
        b = native.frame_bound(len(c))
        caps.append(b)
        ip += native.round16(len(c))
        op += native.round16(b)
    d_in = ctx.device_alloc(max(ip, 16) + 64)
    d_out = ctx.device_alloc(max(op, 16) + 64)
    try:
        for c, o in zip(chunks, src_off):
            if len(c):
                ctx.h2d(d_in + o, c)
        out_lens, digests, ms = ctx.process_device(d_in, src_off, [len(c) for c in chunks], d_out, dst_off, caps, flags, stream)
        frames: List[bytes] = []
        if flags == 0 or (flags & native.F_LZ4):
            frames = [ctx.d2h(d_out + o, n) for o, n in zip(dst_off, out_lens)]
        return frames, digests, out_lens, ms
    finally:
        ctx.device_free(d_in)
        ctx.device_free(d_out)
