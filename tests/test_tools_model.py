"""tools/lz4_window_model.c (the sequential CPU model of the GPU compressor's parse) must emit valid LZ4 blocks:
each block is decoded with the strict oracle decoder wrapped in a minimal frame."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from skyplane_b200 import synth

ROOT = Path(__file__).resolve().parent.parent


class Opts(ctypes.Structure):
    _fields_ = [("hash_log", ctypes.c_int), ("window", ctypes.c_int), ("skip_trigger", ctypes.c_int), ("in_window", ctypes.c_int),
                ("back_ext", ctypes.c_int)]


@pytest.fixture(scope="module")
def model():
    so = ROOT / "tools" / "bin" / "liblz4model.so"
    so.parent.mkdir(exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", str(so), str(ROOT / "tools" / "lz4_window_model.c")])
    m = ctypes.CDLL(str(so))
    m.model_compress_block.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.POINTER(Opts)]
    m.model_compress_block.restype = ctypes.c_uint32
    return m


def frame_of_blocks(data: bytes, model, o: Opts) -> bytes:
    hdr = bytes([0x04, 0x22, 0x4D, 0x18, 0x68, 0x40]) + len(data).to_bytes(8, "little")
    hdr += bytes([(oracle.xxh32(hdr[4:]) >> 8) & 0xFF])
    out = bytearray(hdr)
    buf = ctypes.create_string_buffer(65536 + 16)
    for pos in range(0, len(data), 65536):
        blk = data[pos : pos + 65536]
        c = model.model_compress_block(blk, len(blk), buf, ctypes.byref(o))
        if c:
            out += c.to_bytes(4, "little") + buf.raw[:c]
        else:
            out += (len(blk) | 0x80000000).to_bytes(4, "little") + blk
    return bytes(out + bytes(4))


@pytest.mark.parametrize("opts", [(12, 32, 6, 1, 1), (12, 32, 6, 1, 2), (11, 16, 6, 0, 0), (13, 64, 5, 1, 1)])
def test_model_emits_valid_lz4(model, opts):
    o = Opts(*opts)
    rng = np.random.default_rng(3)
    datas = [rng.bytes(n) for n in (1, 12, 13, 100, 65536)] + [bytes(70000), (b"abcdefg" * 20000)[:131073],
                                                                synth.silesia_like_chunk(11, 300000), b"x" * 13 + rng.bytes(40) + b"x" * 200]
    for d in datas:
        fr = frame_of_blocks(d, model, o)
        assert oracle.lz4f_decode(fr, len(d)) == d
    # the kernel's parse should stay within 10 % of the reference's ratio on the compressible set
    d = synth.silesia_like_chunk(0, 4 << 20)
    if opts[:2] == (12, 32):
        assert len(d) / len(frame_of_blocks(d, model, o)) >= 0.90 * len(d) / len(oracle.lz4f_compress(d))
