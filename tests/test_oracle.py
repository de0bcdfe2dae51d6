"""Pins the CPU oracle (oracle/skyoracle.c) against the reference's engines and golden vectors.

Reference behaviours under test: lz4.frame.compress / decompress (gateway_operator.py:359,
gateway_receiver.py:196) and hashlib.md5 (s3_interface.py:181-192).
"""
import hashlib

import numpy as np
import pytest

import oracle
import oracle.reflib as ref
from skyplane_b200 import synth

RNG = np.random.default_rng(2024)
LENS = [0, 1, 2, 3, 4, 5, 11, 12, 13, 14, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 1000, 4095, 4096, 65535, 65536, 65537,
        65547, 65548, 131071, 131072, 131073, (1 << 20) - 1, 1 << 20, (1 << 20) + 1]


def _kinds(n):
    yield "random", RNG.bytes(n)
    yield "zeros", bytes(n)
    yield "period7", (b"abcdefg" * (n // 7 + 1))[:n]
    yield "text", (b"it was the best of times, it was the worst of times; " * (n // 50 + 1))[:n]
    yield "half", ((b"lorem ipsum dolor sit amet " * (n // 54 + 1))[: n // 2] + RNG.bytes(n - n // 2))


# ---------------------------------------------------------------- MD5
def test_md5_rfc1321_suite(golden):
    for case in golden["md5_kat"]["rfc1321"]:
        assert oracle.md5(bytes.fromhex(case["msg_hex"])).hex() == case["md5"]


def test_md5_padding_boundaries(golden):
    for case in golden["md5_kat"]["zeros"]:
        assert oracle.md5(bytes(case["len"])).hex() == case["md5"], case["len"]


def test_md5_seeded_golden(golden):
    rng = np.random.default_rng(1321)
    for case in golden["md5_kat"]["seeded"]:
        m = rng.bytes(case["len"])
        if case["msg_hex"] is not None:
            assert m.hex() == case["msg_hex"]
        assert oracle.md5(m).hex() == case["md5"]


@pytest.mark.parametrize("n", LENS + [8 << 20])
def test_md5_matches_hashlib(n):
    m = RNG.bytes(n)
    assert oracle.md5(m) == hashlib.md5(m).digest()


# ---------------------------------------------------------------- XXH32
def test_xxh32_matches_xxhash_module():
    xxhash = pytest.importorskip("xxhash")
    for n in [0, 1, 2, 3, 4, 9, 10, 15, 16, 17, 31, 32, 33, 1000]:
        m = RNG.bytes(n)
        for seed in (0, 1, 0x9E3779B1):
            assert oracle.xxh32(m, seed) == xxhash.xxh32(m, seed=seed).intdigest()


def test_frame_header_checksum_values():
    # FLG 0x68 / 0x48, BD 0x40, content size 8 MiB -> 0x3F / 0x2D
    size = (8 << 20).to_bytes(8, "little")
    assert (oracle.xxh32(bytes([0x68, 0x40]) + size) >> 8) & 0xFF == 0x3F
    assert (oracle.xxh32(bytes([0x48, 0x40]) + size) >> 8) & 0xFF == 0x2D


# ---------------------------------------------------------------- LZ4 frame compressor: byte-exact vs liblz4
@pytest.mark.skipif(not ref.available(), reason="liblz4.so.1 not present")
@pytest.mark.parametrize("n", LENS)
def test_compress_is_byte_identical_to_liblz4(n):
    for kind, data in _kinds(n):
        assert oracle.lz4f_compress(data) == ref.lz4f_compress(data), (kind, n)


@pytest.mark.skipif(not ref.available(), reason="liblz4.so.1 not present")
def test_compress_byte_identical_on_bench_workloads():
    for data in (synth.random_chunk(0, 8 << 20), synth.silesia_like_chunk(0, 16 << 20), synth.silesia_like_chunk(5, (4 << 20) + 12345)):
        assert oracle.lz4f_compress(data) == ref.lz4f_compress(data)


def test_compress_golden_frames(golden):
    g = golden["lz4_frames"]
    for case in g["cases"]:
        data = bytes.fromhex(case["input_hex"])
        frame = bytes.fromhex(case["frame_hex"])
        assert oracle.lz4f_compress(data) == frame, case["name"]
        assert oracle.lz4f_decode(frame, len(data)) == data, case["name"]
        assert oracle.md5(data).hex() == case["md5"]
    for case in g["big"]:
        n = case["len"]
        data = bytes(n) if case["seed"] is None else np.random.default_rng(case["seed"]).bytes(n)
        fr = oracle.lz4f_compress(data)
        assert len(fr) == case["frame_len"] and hashlib.md5(fr).hexdigest() == case["frame_md5"], case["name"]
        assert fr[:19].hex() == case["frame_head_hex"]
        assert oracle.md5(data).hex() == case["md5"]


def test_frame_bound_matches_liblz4_worst_case():
    for n in (1, 100, 65536, 65537, 1 << 20, 8 << 20):
        data = np.random.default_rng(n).bytes(n)
        assert len(oracle.lz4f_compress(data)) <= oracle.lz4f_bound(n)
    assert oracle.lz4f_bound(8 << 20) == 8389139  # what liblz4 emits for 8 MiB of random bytes


# ---------------------------------------------------------------- LZ4 frame decoder
@pytest.mark.skipif(not ref.available(), reason="liblz4.so.1 not present")
@pytest.mark.parametrize("n", [0, 1, 13, 100, 65536, 65537, 200000, 1 << 20])
def test_decoder_agrees_with_liblz4_and_pyarrow(n):
    pa = pytest.importorskip("pyarrow")
    for kind, data in _kinds(n):
        for frame in (ref.lz4f_compress(data), oracle.lz4f_compress_indep(data)):
            assert oracle.lz4f_decode(frame, n) == data, kind
            assert ref.lz4f_decompress(frame, n) == data, kind
            if n:
                assert pa.decompress(frame, decompressed_size=n, codec="lz4").to_pybytes() == data, kind


def test_decoder_rejects_malformed_frames():
    data = (b"hello hello hello hello hello hello " * 100)
    good = oracle.lz4f_compress(data)
    assert oracle.lz4f_decode(good) == data
    with pytest.raises(oracle.OracleError):  # bad magic
        oracle.lz4f_decode(b"\x00" + good[1:], len(data))
    with pytest.raises(oracle.OracleError):  # header checksum
        oracle.lz4f_decode(good[:14] + bytes([good[14] ^ 1]) + good[15:], len(data))
    with pytest.raises(oracle.OracleError):  # truncated
        oracle.lz4f_decode(good[:-3], len(data))
    with pytest.raises(oracle.OracleError):  # wrong content size
        bad = bytearray(good)
        bad[6] ^= 1
        bad[14] = (oracle.xxh32(bytes(bad[4:14])) >> 8) & 0xFF
        oracle.lz4f_decode(bytes(bad), len(data) + 8)
    # a "compressed" block larger than the 64 KiB maximum must be refused
    hdr = good[:15]
    big = hdr + (70000).to_bytes(4, "little") + bytes(70000) + bytes(4)
    with pytest.raises(oracle.OracleError) as e:
        oracle.lz4f_decode(big, 1 << 20)
    assert e.value.code == -4
    # offset pointing before the start of an independent block
    indep = bytearray(oracle.lz4f_compress_indep(bytes(200000)))
    info = oracle.lz4f_decode(bytes(indep), 200000, with_info=True)[1]
    assert info["flg"] == 0x68 and info["blocks"] == 4


def test_indep_layout_properties():
    data = synth.silesia_like_chunk(3, 1 << 20)
    fr = oracle.lz4f_compress_indep(data)
    out, info = oracle.lz4f_decode(fr, len(data), with_info=True)
    assert out == data and info["flg"] == 0x68 and info["bd"] == 0x40 and info["blocks"] == 16
    rnd = synth.random_chunk(1, 1 << 20)
    fr = oracle.lz4f_compress_indep(rnd)
    out, info = oracle.lz4f_decode(fr, len(rnd), with_info=True)
    assert out == rnd and info["raw_blocks"] == 16 and len(fr) == oracle.lz4f_bound(len(rnd))


def test_chunk_stage_pair():
    data = synth.silesia_like_chunk(1, 300000)
    frame, digest = oracle.chunk_stage(data)
    assert digest == hashlib.md5(data).digest()
    assert oracle.lz4f_decode(frame, len(data)) == data


# ---------------------------------------------------------------- property tests (hypothesis): structured random inputs
def _structured(draw_bytes, rng_seed, n_segments):
    """Concatenation of segments: random bytes, runs, repeats of earlier content, ascii words."""
    r = np.random.default_rng(rng_seed)
    out = bytearray()
    for _ in range(n_segments):
        kind = int(r.integers(0, 5))
        ln = int(r.integers(1, 3000))
        if kind == 0:
            out += r.bytes(ln)
        elif kind == 1:
            out += bytes([int(r.integers(0, 256))]) * ln
        elif kind == 2 and len(out) > 8:
            start = int(r.integers(0, len(out) - 4))
            seg = bytes(out[start : start + min(ln, len(out) - start)])
            out += seg * int(r.integers(1, 4))
        elif kind == 3:
            out += (b"alpha beta gamma delta " * (ln // 23 + 1))[:ln]
        else:
            per = int(r.integers(1, 40))
            out += (r.bytes(per) * (ln // per + 1))[:ln]
    return bytes(out)


try:
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @pytest.mark.skipif(not ref.available(), reason="liblz4.so.1 not present")
    @settings(max_examples=120, deadline=None)
    @given(seed=st.integers(0, 2**32 - 1), nseg=st.integers(0, 60), pad=st.integers(0, 70000))
    def test_property_compressor_byte_identical_to_liblz4(seed, nseg, pad):
        data = _structured(None, seed, nseg) + bytes(pad % 7) + np.random.default_rng(seed ^ 1).bytes(pad if seed % 3 == 0 else 0)
        frame = oracle.lz4f_compress(data)
        assert frame == ref.lz4f_compress(data)
        assert oracle.lz4f_decode(frame, len(data)) == data
        indep = oracle.lz4f_compress_indep(data)
        assert oracle.lz4f_decode(indep, len(data)) == data == ref.lz4f_decompress(indep, len(data))
        assert oracle.md5(data) == hashlib.md5(data).digest()

    @settings(max_examples=200, deadline=None)
    @given(blob=st.binary(min_size=0, max_size=400), raw=st.integers(0, 5000))
    def test_property_decoder_never_crashes_on_garbage(blob, raw):
        """Arbitrary bytes either decode (and then liblz4 agrees) or raise OracleError -- no crash, no overrun."""
        hdr = bytes.fromhex("04224d186040") + bytes([(oracle.xxh32(bytes([0x60, 0x40])) >> 8) & 0xFF])
        frame = hdr + blob
        try:
            out = oracle.lz4f_decode(frame, raw)
        except oracle.OracleError:
            return
        if ref.available():
            assert ref.lz4f_decompress(frame[: len(frame)], max(raw, 1))[: len(out)] == out
except ImportError:  # hypothesis not installed: the parametrised differential tests above still run
    pass


# ---------------------------------------------------------------- XSalsa20-Poly1305 groundwork
def test_secretbox_matches_pynacl():
    """The reference encrypts the compressed frame with nacl.secret.SecretBox (gateway_operator.py:362-364); PyNaCl is
    installed here, so the restatement is pinned against the reference's own engine."""
    nacl_secret = pytest.importorskip("nacl.secret")
    rng = np.random.default_rng(8)
    for n in [0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 1000, 4096, 65537, 300001]:
        key, nonce, msg = rng.bytes(32), rng.bytes(24), rng.bytes(n)
        box = nacl_secret.SecretBox(key)
        want = bytes(box.encrypt(msg, nonce))  # nonce || tag || ciphertext
        got = oracle.secretbox_seal(key, nonce, msg)
        assert nonce + got == want, n
        assert oracle.secretbox_open(key, nonce, got) == msg == box.decrypt(want)
        if n:
            bad = bytearray(got)
            bad[-1] ^= 1
            with pytest.raises(ValueError):
                oracle.secretbox_open(key, nonce, bytes(bad))
