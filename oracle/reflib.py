"""The reference's own CPU engine for this path, bound without python-lz4 -- TEST INFRASTRUCTURE ONLY.

The reference calls ``lz4.frame.compress(data)`` (skyplane/gateway/operators/gateway_operator.py:359),
``lz4.frame.decompress`` (skyplane/gateway/operators/gateway_receiver.py:196) and ``hashlib.md5``
(skyplane/obj_store/s3_interface.py:181-192).  python-lz4 (pinned 4.3.2, poetry.lock:1540) is not
installed in this image, but the library it wraps is: the system ``liblz4.so.1`` (1.9.4).  This module
drives ``LZ4F_compressFrame`` through ctypes with exactly the preferences python-lz4's ``compress()``
fills in for its defaults (block_size=0 -> 64 KiB, block_linked=True, compression_level=0,
content_checksum=False, block_checksum=False, store_size=True, autoFlush=0), so its output is what
the reference would put on the wire.  ``hashlib`` is used directly -- it IS the reference's call.
"""
from __future__ import annotations

import ctypes
import hashlib


class _FrameInfo(ctypes.Structure):
    _fields_ = [
        ("blockSizeID", ctypes.c_int),
        ("blockMode", ctypes.c_int),
        ("contentChecksumFlag", ctypes.c_int),
        ("frameType", ctypes.c_int),
        ("contentSize", ctypes.c_ulonglong),
        ("dictID", ctypes.c_uint),
        ("blockChecksumFlag", ctypes.c_int),
    ]


class _Prefs(ctypes.Structure):
    _fields_ = [
        ("frameInfo", _FrameInfo),
        ("compressionLevel", ctypes.c_int),
        ("autoFlush", ctypes.c_uint),
        ("favorDecSpeed", ctypes.c_uint),
        ("reserved", ctypes.c_uint * 3),
    ]


class _DecOpts(ctypes.Structure):
    _fields_ = [("stableDst", ctypes.c_uint), ("skipChecksums", ctypes.c_uint), ("reserved", ctypes.c_uint * 2)]


_L = None


def available() -> bool:
    try:
        _lib()
        return True
    except OSError:
        return False


def _lib():
    global _L
    if _L is None:
        L = ctypes.CDLL("liblz4.so.1")
        L.LZ4_versionString.restype = ctypes.c_char_p
        L.LZ4F_compressFrameBound.argtypes = [ctypes.c_size_t, ctypes.POINTER(_Prefs)]
        L.LZ4F_compressFrameBound.restype = ctypes.c_size_t
        L.LZ4F_compressFrame.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(_Prefs)]
        L.LZ4F_compressFrame.restype = ctypes.c_size_t
        L.LZ4F_isError.argtypes = [ctypes.c_size_t]
        L.LZ4F_isError.restype = ctypes.c_uint
        L.LZ4F_getErrorName.argtypes = [ctypes.c_size_t]
        L.LZ4F_getErrorName.restype = ctypes.c_char_p
        L.LZ4F_createDecompressionContext.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint]
        L.LZ4F_createDecompressionContext.restype = ctypes.c_size_t
        L.LZ4F_freeDecompressionContext.argtypes = [ctypes.c_void_p]
        L.LZ4F_freeDecompressionContext.restype = ctypes.c_size_t
        L.LZ4F_decompress.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t),
            ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(_DecOpts),
        ]
        L.LZ4F_decompress.restype = ctypes.c_size_t
        _L = L
    return _L


def version() -> str:
    return _lib().LZ4_versionString().decode()


def _prefs(n: int) -> _Prefs:
    p = _Prefs()
    p.frameInfo.blockSizeID = 0  # LZ4F_default (64 KiB)
    p.frameInfo.blockMode = 0  # LZ4F_blockLinked
    p.frameInfo.contentSize = n  # store_size=True
    p.compressionLevel = 0
    p.autoFlush = 0
    return p


def _addr(buf):
    mv = memoryview(buf).cast("B")
    if mv.nbytes == 0:
        return None, 0, mv
    if isinstance(buf, bytes):
        return ctypes.cast(ctypes.c_char_p(buf), ctypes.c_void_p), mv.nbytes, buf
    if mv.readonly:
        arr = (ctypes.c_char * mv.nbytes).from_buffer_copy(mv)
    else:
        arr = (ctypes.c_char * mv.nbytes).from_buffer(mv)
    return ctypes.cast(arr, ctypes.c_void_p), mv.nbytes, arr


def lz4f_compress(data) -> bytes:
    """``lz4.frame.compress(data)`` with python-lz4's defaults, via liblz4.so.1."""
    L = _lib()
    p, n, _k = _addr(data)
    prefs = _prefs(n)
    cap = L.LZ4F_compressFrameBound(n, ctypes.byref(prefs))
    out = ctypes.create_string_buffer(cap)
    r = L.LZ4F_compressFrame(ctypes.cast(out, ctypes.c_void_p), cap, p, n, ctypes.byref(prefs))
    if L.LZ4F_isError(r):
        raise ValueError(L.LZ4F_getErrorName(r).decode())
    return out.raw[:r]


class Compressor:
    """Re-usable buffers for timing loops (no allocation inside the timed region)."""

    def __init__(self, max_n: int):
        self.L = _lib()
        self.prefs = _prefs(max_n)
        self.cap = self.L.LZ4F_compressFrameBound(max_n, ctypes.byref(self.prefs))
        self.out = ctypes.create_string_buffer(self.cap)

    def compress_into(self, addr: int, n: int) -> int:
        self.prefs.frameInfo.contentSize = n
        r = self.L.LZ4F_compressFrame(ctypes.cast(self.out, ctypes.c_void_p), self.cap, addr, n, ctypes.byref(self.prefs))
        if self.L.LZ4F_isError(r):
            raise ValueError(self.L.LZ4F_getErrorName(r).decode())
        return r


def lz4f_decompress(frame, max_out: int) -> bytes:
    """``lz4.frame.decompress(frame)`` via liblz4.so.1's streaming decoder. Raises ValueError on bad frames."""
    L = _lib()
    ctx = ctypes.c_void_p()
    r = L.LZ4F_createDecompressionContext(ctypes.byref(ctx), 100)
    if L.LZ4F_isError(r):
        raise ValueError(L.LZ4F_getErrorName(r).decode())
    try:
        p, n, _k = _addr(frame)
        out = ctypes.create_string_buffer(max(1, max_out))
        src_off, dst_off = 0, 0
        hint = 1
        base_src = p.value if p is not None else 0
        base_dst = ctypes.addressof(out)
        while hint != 0:
            s = ctypes.c_size_t(n - src_off)
            d = ctypes.c_size_t(max_out - dst_off)
            hint = L.LZ4F_decompress(ctx, base_dst + dst_off, ctypes.byref(d), base_src + src_off, ctypes.byref(s), None)
            if L.LZ4F_isError(hint):
                raise ValueError(L.LZ4F_getErrorName(hint).decode())
            src_off += s.value
            dst_off += d.value
            if hint != 0 and s.value == 0 and d.value == 0:
                raise ValueError("truncated frame or destination too small")
        return out.raw[:dst_off]
    finally:
        L.LZ4F_freeDecompressionContext(ctx)


def md5(data) -> bytes:
    return hashlib.md5(data).digest()
