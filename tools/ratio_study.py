#!/usr/bin/env python
"""Compression-ratio study on the CPU model of the GPU block compressor (tools/lz4_window_model.c).
Prints frame-size ratios for design variants next to liblz4's (linked blocks, what the reference emits).
Every model output is decoded with liblz4 to prove the variant still emits valid LZ4."""
import ctypes
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from skyplane_b200 import synth  # noqa: E402

SO = ROOT / "tools" / "bin" / "liblz4model.so"
SO.parent.mkdir(exist_ok=True)
subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(SO), str(ROOT / "tools" / "lz4_window_model.c")])
M = ctypes.CDLL(str(SO))
LZ4 = ctypes.CDLL("liblz4.so.1")


class Opts(ctypes.Structure):
    _fields_ = [("hash_log", ctypes.c_int), ("window", ctypes.c_int), ("skip_trigger", ctypes.c_int), ("in_window", ctypes.c_int),
                ("back_ext", ctypes.c_int)]


M.model_compress_block.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.POINTER(Opts)]
M.model_compress_block.restype = ctypes.c_uint32
LZ4.LZ4_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
LZ4.LZ4_compress_default.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]


def model_size(data: bytes, o: Opts, check: bool = True) -> int:
    total = 15 + 4
    out = ctypes.create_string_buffer(65536 + 16)
    back = ctypes.create_string_buffer(65536)
    for pos in range(0, len(data), 65536):
        blk = data[pos : pos + 65536]
        c = M.model_compress_block(blk, len(blk), out, ctypes.byref(o))
        if c and check:
            r = LZ4.LZ4_decompress_safe(out, back, c, len(blk))
            assert r == len(blk) and back.raw[:r] == blk, "model emitted an invalid block"
        total += 4 + (c or len(blk))
    return total


def liblz4_linked_size(data: bytes) -> int:
    sys.path.insert(0, str(ROOT))
    import oracle.reflib as ref  # dev tool: only used to print the reference's ratio next to ours

    return len(ref.lz4f_compress(data))


VARIANTS = {
    "kernel (h12 w32 inwin backext>1)": Opts(12, 32, 6, 1, 1),
    "no in-window candidates": Opts(12, 32, 6, 0, 1),
    "backward extension always": Opts(12, 32, 6, 1, 2),
    "no backward extension": Opts(12, 32, 6, 1, 0),
    "hash_log 11 (4 KiB table)": Opts(11, 32, 6, 1, 1),
    "hash_log 13 (16 KiB table)": Opts(13, 32, 6, 1, 1),
    "window 16": Opts(12, 16, 6, 1, 1),
    "window 64": Opts(12, 64, 6, 1, 1),
    "skip trigger 5": Opts(12, 32, 5, 1, 1),
    "skip trigger 7": Opts(12, 32, 7, 1, 1),
}

if __name__ == "__main__":
    sets = {"silesia-like 4 x 4 MiB": [synth.silesia_like_chunk(i, 4 << 20) for i in range(4)],
            "text-only 4 MiB": [synth._text(__import__("numpy").random.default_rng(5), 4 << 20)],
            "records 4 MiB": [synth._records(__import__("numpy").random.default_rng(6), 4 << 20)],
            "numeric 4 MiB": [synth._numeric(__import__("numpy").random.default_rng(7), 4 << 20)]}
    rows = []
    for sname, datas in sets.items():
        raw = sum(map(len, datas))
        refsz = sum(liblz4_linked_size(d) for d in datas)
        row = {"set": sname, "reference (liblz4 linked)": round(raw / refsz, 4)}
        for vname, o in VARIANTS.items():
            row[vname] = round(raw / sum(model_size(d, o) for d in datas), 4)
        rows.append(row)
        print(json.dumps(row), flush=True)
