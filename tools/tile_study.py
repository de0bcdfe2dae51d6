#!/usr/bin/env python
"""Compression-ratio study on the CPU twin of the round-2 GPU block compressor (tools/lz4_tile_model.c).
Prints frame-size ratios for design variants next to liblz4's (linked blocks = what the reference emits).
Every model output is decoded with liblz4 to prove the variant emits valid LZ4."""
import ctypes
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skyplane_b200 import synth  # noqa: E402

SO = ROOT / "tools" / "bin" / "liblz4tile.so"
SO.parent.mkdir(exist_ok=True)
subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(SO), str(ROOT / "tools" / "lz4_tile_model.c")])
M = ctypes.CDLL(str(SO))
LZ4 = ctypes.CDLL("liblz4.so.1")


class Opts(ctypes.Structure):
    _fields_ = [("entries", ctypes.c_int), ("tag_bits", ctypes.c_int), ("tile", ctypes.c_int), ("max_step", ctypes.c_int),
                ("back_ext", ctypes.c_int), ("ways", ctypes.c_int), ("policy", ctypes.c_int), ("hash5", ctypes.c_int)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("probes", "hits", "verified", "accepted", "tiles")]


M.tile_compress_block.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.POINTER(Opts)]
M.tile_compress_block.restype = ctypes.c_uint32
LZ4.LZ4_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]


def model_size(data: bytes, o: Opts, check: bool = True) -> int:
    total = 15 + 4
    out = ctypes.create_string_buffer(65536 + 16)
    back = ctypes.create_string_buffer(65536)
    for pos in range(0, len(data), 65536):
        blk = data[pos: pos + 65536]
        c = M.tile_compress_block(blk, len(blk), out, ctypes.byref(o))
        if c and check:
            r = LZ4.LZ4_decompress_safe(out, back, c, len(blk))
            assert r == len(blk) and back.raw[:r] == blk, "model emitted an invalid block"
        total += 4 + (c or len(blk))
    return total


def liblz4_linked_size(data: bytes) -> int:
    import oracle.reflib as ref  # dev tool: only used to print the reference's ratio next to ours

    return len(ref.lz4f_compress(data))


VARIANTS = {
    "h5m p0 back2 4096e T256 step16": Opts(4096, 16, 256, 16, 2, 1, 0, 2),
    "h5m p0 back2 3072e T256 step16": Opts(3072, 16, 256, 16, 2, 1, 0, 2),
    "h5m p0 back2 2048e T256 step16": Opts(2048, 16, 256, 16, 2, 1, 0, 2),
    "h5m p0 back2 4096e T256 step1": Opts(4096, 16, 256, 1, 2, 1, 0, 2),
}

if __name__ == "__main__":
    sets = {"silesia-like 4 x 4 MiB": [synth.silesia_like_chunk(i, 4 << 20) for i in range(4)],
            "text-only 4 MiB": [synth._text(np.random.default_rng(5), 4 << 20)],
            "records 4 MiB": [synth._records(np.random.default_rng(6), 4 << 20)],
            "numeric 4 MiB": [synth._numeric(np.random.default_rng(7), 4 << 20)],
            "random 1 MiB": [synth.random_chunk(0, 1 << 20)]}
    for sname, datas in sets.items():
        raw = sum(map(len, datas))
        refsz = sum(liblz4_linked_size(d) for d in datas)
        row = {"set": sname, "reference (liblz4 linked)": round(raw / refsz, 4)}
        for vname, o in VARIANTS.items():
            st = Stats()
            M.tile_model_stats(ctypes.byref(st), 1)
            row[vname] = round(raw / sum(model_size(d, o) for d in datas), 4)
            M.tile_model_stats(ctypes.byref(st), 1)
            if vname.startswith("h5m p0 back2 4096e T256 step16"):
                row["stats"] = {"probes/B": round(st.probes / raw, 3), "hits/B": round(st.hits / raw, 3),
                                "verified/B": round(st.verified / raw, 4), "B/seq": round(raw / max(1, st.accepted), 1)}
        print(json.dumps(row), flush=True)
