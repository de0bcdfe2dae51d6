"""GPU parity tests of the receiver-side stage: LZ4 frame decode + MD5 of the decoded
bytes, against the reference's decoder (liblz4 LZ4F_decompress = lz4.frame.decompress, gateway_receiver.py:196),
the strict oracle decoder and hashlib.  Frames come from three encoders: the reference's (liblz4, LINKED blocks),
the oracle's independent-block variant, and the GPU stage itself."""
import hashlib
import multiprocessing as mp

import numpy as np
import pytest

import oracle
import oracle.reflib as ref
from gpu_util import run_device
from skyplane_b200 import native, synth
from skyplane_b200.chunk import Chunk, ChunkRequest
from skyplane_b200.chunk_store import ChunkStore
from skyplane_b200.gateway_queue import GatewayQueue
from skyplane_b200.operators import ChecksumMismatchException, GatewayCompressHash, GatewayDecompressVerify
from skyplane_b200.stage import ChunkStage

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]
RNG = np.random.default_rng(99)
LENS = [0, 1, 12, 13, 100, 4096, 65535, 65536, 65537, 131073, 200000, (1 << 20) + 5]


@pytest.fixture(scope="module")
def ctx():
    c = native.Context(0, 1 << 30, 4096, 0)
    yield c
    c.close()


def decode_device(ctx, frames, raw_lens):
    f_off, o_off, fp, op = [], [], 0, 0
    for f, r in zip(frames, raw_lens):
        f_off.append(fp)
        o_off.append(op)
        fp += native.round16(len(f))
        op += native.round16(r)
    d_f = ctx.device_alloc(fp + 64)
    d_o = ctx.device_alloc(op + 64)
    try:
        for f, o in zip(frames, f_off):
            ctx.h2d(d_f + o, f)
        st, dg, ms = ctx.decode_device(d_f, f_off, [len(f) for f in frames], d_o, o_off, raw_lens)
        outs = [ctx.d2h(d_o + o, r) if s == 0 else None for o, r, s in zip(o_off, raw_lens, st)]
        return outs, dg, st
    finally:
        ctx.device_free(d_f)
        ctx.device_free(d_o)


def kinds(n):
    return [RNG.bytes(n), bytes(n), (b"abcdefg" * (n // 7 + 1))[:n], (b"the quick brown fox jumps over the lazy dog " * (n // 40 + 1))[:n],
            (b"lorem ipsum dolor " * (n // 36 + 1))[: n // 2] + RNG.bytes(n - n // 2)]


@pytest.mark.parametrize("encoder", ["reference_linked", "oracle_indep", "gpu"])
def test_decode_matches_reference_decoder(ctx, encoder):
    datas = [d for n in LENS for d in kinds(n)] + [synth.silesia_like_chunk(40, 3 << 20), synth.random_chunk(41, 2 << 20)]
    if encoder == "reference_linked":
        frames = [ref.lz4f_compress(d) for d in datas]  # what the reference's CPU sender puts on the wire
    elif encoder == "oracle_indep":
        frames = [oracle.lz4f_compress_indep(d) for d in datas]
    else:
        frames = run_device(ctx, datas)[0]
    outs, digests, status = decode_device(ctx, frames, [len(d) for d in datas])
    for d, f, o, dg, st in zip(datas, frames, outs, digests, status):
        assert st == 0, (encoder, len(d), st)
        assert o == d == ref.lz4f_decompress(f, len(d)) == oracle.lz4f_decode(f, len(d))
        assert dg == hashlib.md5(d).digest()


def test_decode_rejects_malformed_frames_without_crashing(ctx):
    data = synth.silesia_like_chunk(42, 300000)
    good = oracle.lz4f_compress_indep(data)
    n = len(data)
    bad_magic = b"\x00" + good[1:]
    bad_hc = good[:14] + bytes([good[14] ^ 1]) + good[15:]
    truncated = good[:-9]
    wrong_size = good  # decoded against raw_len + 1
    # a sequence whose match offset is zero: zeros compress to token 0x1F, one literal, offset 1, ...
    zgood = oracle.lz4f_compress_indep(bytes(n))
    assert zgood[19] >> 4 == 1 and zgood[21:23] == b"\x01\x00"
    zero_offset = zgood[:21] + b"\x00\x00" + zgood[23:]
    flipped = bytearray(good)
    for k in range(40, len(flipped) - 8, 97):
        flipped[k] ^= 0x5A
    frames = [good, bad_magic, bad_hc, truncated, wrong_size, zero_offset, bytes(flipped)]
    raws = [n, n, n, n, n + 1, n, n]
    outs, digests, status = decode_device(ctx, frames, raws)
    assert status[0] == 0 and outs[0] == data
    assert status[1] == native.D_BAD_HEADER and status[2] == native.D_BAD_HEADER
    assert status[3] == native.D_TRUNCATED and status[4] == native.D_SIZE
    assert status[5] == native.D_CORRUPT
    for f, r, st, o in zip(frames[5:], raws[5:], status[5:], outs[5:]):
        try:
            want = oracle.lz4f_decode(f, r)
        except oracle.OracleError:
            want = None
        if want is None:
            assert st != 0  # the strict oracle rejects it: so must we
        else:
            assert st != 0 or o == want  # a flip that still parses must decode to the same bytes as the oracle


def test_stage_roundtrip_and_host_decode():
    stage = ChunkStage(0, max_batch_bytes=64 << 20, max_chunks=32, n_slots=1)
    try:
        datas = [synth.silesia_like_chunk(50 + i, (2 << 20) + i * 7919) for i in range(6)] + [b"", b"x", synth.random_chunk(7, 1 << 20)]
        res = stage.process(datas)
        back = stage.decode([bytes(r.frame) for r in res], [len(d) for d in datas])
        for d, r, (o, dg, st) in zip(datas, res, back):
            assert st == 0 and o == d and dg == r.md5 == hashlib.md5(d).digest()
        # the reference sender's frames (linked blocks) through the host path
        back = stage.decode([ref.lz4f_compress(d) for d in datas], [len(d) for d in datas])
        assert all(st == 0 and o == d for d, (o, dg, st) in zip(datas, back))
    finally:
        stage.close()


def test_sender_and_receiver_operators_end_to_end(tmp_path):
    """compress_hash on the source gateway -> (wire) -> decompress_verify on the destination gateway."""
    src, dst = ChunkStore(tmp_path / "src"), ChunkStore(tmp_path / "dst")
    ev, eq = mp.Event(), mp.Queue()
    comp = GatewayCompressHash("ch", "r", GatewayQueue(), None, ev, eq, src)
    dec = GatewayDecompressVerify("dv", "r", GatewayQueue(), None, ev, eq, dst)
    comp.worker_id = dec.worker_id = 0
    try:
        data = synth.silesia_like_chunk(60, (5 << 20) + 321)
        cid = "ab" * 16
        src.get_chunk_file_path(cid).write_bytes(data)
        req = ChunkRequest(Chunk("k", "k", cid, len(data), partition_id="0"))
        assert comp.process(req) is True and req.chunk.md5_hash == hashlib.md5(data).digest()
        assert dec.process(req) is False  # payload has not arrived at the destination yet
        dst.get_compressed_file_path(cid).write_bytes(src.get_compressed_file_path(cid).read_bytes())  # "the wire"
        assert dec.process(req) is True
        assert dst.get_chunk_file_path(cid).read_bytes() == data and not dst.get_compressed_file_path(cid).exists()
        # tampered digest -> ChecksumMismatchException ; tampered payload -> rejected frame or checksum mismatch
        dst.get_compressed_file_path(cid).write_bytes(src.get_compressed_file_path(cid).read_bytes())
        req.chunk.md5_hash = bytes(16)
        with pytest.raises(ChecksumMismatchException):
            dec.process(req)
        frame = bytearray(src.get_compressed_file_path(cid).read_bytes())
        frame[len(frame) // 2] ^= 0xFF
        dst.get_compressed_file_path(cid).write_bytes(bytes(frame))
        req.chunk.md5_hash = hashlib.md5(data).digest()
        with pytest.raises((ChecksumMismatchException, ValueError)):
            dec.process(req)
    finally:
        comp.worker_exit(0)
        dec.worker_exit(0)
