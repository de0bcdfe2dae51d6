"""GPU tests of the rows next to the hot path through the C ABI:
  * XSalsa20-Poly1305 on the device: boxes byte-identical to PyNaCl's SecretBox (the reference's engine,
    gateway_operator.py:183-186, :362-364) and to the oracle; open + tamper detection (gateway_receiver.py:191-193);
  * `compress: false` (gateway_daemon.py:235): digest only, payload passes through;
  * sender hand-off without a frame file: GPU stage -> wire framing from the pinned slot -> socket -> receive into the
    receiver stage's pinned buffer -> GPU decode + digest (gateway_operator.py:367-402, gateway_receiver.py:150-233)."""
import hashlib
import os
import socket
import threading

import numpy as np
import pytest

import oracle
from skyplane_b200 import native, synth, wire
from skyplane_b200.chunk import Chunk
from skyplane_b200.stage import ChunkStage

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

KEY = bytes((7 * i + 3) & 0xFF for i in range(32))
LENS = [0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 4095, 4096, 4097, 65536, 100000, (1 << 20) + 5]


@pytest.fixture(scope="module")
def stage():
    s = ChunkStage(0, max_batch_bytes=64 << 20, max_chunks=64, n_slots=2)
    s.set_e2ee_key(KEY)
    yield s
    s.close()


def test_boxes_equal_pynacl_and_oracle(stage):
    """compress=False + encrypt=True seals the raw chunk: the box must be SecretBox(key).encrypt(chunk, nonce) byte for byte."""
    nacl_secret = pytest.importorskip("nacl.secret")
    rng = np.random.default_rng(5)
    datas = [rng.bytes(n) for n in LENS]
    nonces = rng.bytes(24 * len(datas))
    res = stage.process(datas, compress=False, encrypt=True, nonces=nonces)
    box = nacl_secret.SecretBox(KEY)
    for i, (d, r) in enumerate(zip(datas, res)):
        nonce = nonces[24 * i: 24 * i + 24]
        want = bytes(box.encrypt(d, nonce))  # nonce | tag | ciphertext
        assert bytes(r.frame) == want, f"len {len(d)}"
        assert want[24:] == oracle.secretbox_seal(KEY, nonce, d)
        assert r.md5 == hashlib.md5(d).digest() and r.is_encrypted and not r.is_compressed and r.comp_len == len(d) + 40


def test_sealed_frames_open_with_pynacl_and_on_the_gpu(stage):
    nacl_secret = pytest.importorskip("nacl.secret")
    datas = [synth.silesia_like_chunk(30 + i, 300000 + 4321 * i) for i in range(4)] + [synth.random_chunk(9, 70000), b"", b"tiny"]
    res = stage.process(datas, compress=True, encrypt=True)
    box = nacl_secret.SecretBox(KEY)
    for d, r in zip(datas, res):
        frame = box.decrypt(bytes(r.frame))  # what the destination gateway does (gateway_receiver.py:193)
        assert oracle.lz4f_decode(frame, len(d)) == d and r.is_compressed and r.is_encrypted
    out = stage.decode([bytes(r.frame) for r in res], [len(d) for d in datas], encrypted=True)
    for d, (data, dg, st) in zip(datas, out):
        assert st == 0 and data == d and dg == hashlib.md5(d).digest()
    # tampering (ciphertext, tag, nonce) and truncation are authentication failures, and no bytes come back
    good = bytes(res[0].frame)
    bad = [good[:50] + bytes([good[50] ^ 1]) + good[51:], good[:30] + bytes([good[30] ^ 0x80]) + good[31:],
           bytes([good[0] ^ 1]) + good[1:], good[:-1], good[:39]]
    out = stage.decode(bad, [len(datas[0])] * len(bad), encrypted=True)
    assert all(st == native.D_AUTH and data is None for data, _, st in out)
    # a box sealed by PyNaCl opens on the GPU too
    theirs = bytes(box.encrypt(oracle.lz4f_compress_indep(datas[1])))
    (data, dg, st), = stage.decode([theirs], [len(datas[1])], encrypted=True)
    assert st == 0 and data == datas[1]


def test_e2ee_needs_a_key():
    s = ChunkStage(0, max_batch_bytes=1 << 20, max_chunks=4, n_slots=1)
    try:
        with pytest.raises(native.SkyChunkError) as e:
            s.process([b"abc"], encrypt=True)
        assert e.value.code == native.SKY_E_NOKEY
    finally:
        s.close()


def test_compress_false_is_digest_only(stage):
    datas = [synth.random_chunk(1, 200000), synth.silesia_like_chunk(2, 70000), b""]
    res = stage.process(datas, compress=False)
    for d, r in zip(datas, res):
        assert bytes(r.frame) == d and r.md5 == hashlib.md5(d).digest() and not r.is_compressed and r.comp_len == r.raw_len == len(d)


@pytest.mark.parametrize("encrypt", [False, True])
def test_sender_sink_to_receiver_stage_over_a_socket(stage, encrypt):
    """f2: payloads leave the sender from the pinned slot (no frame file), arrive in the receiver's pinned staging buffer,
    and are opened / decoded / digested on the GPU."""
    datas = [synth.silesia_like_chunk(50 + i, (2 << 20) + 999 * i) for i in range(3)] + [synth.random_chunk(3, 1 << 20), b"", b"z" * 13]
    chunks = [Chunk("src", "dst", "%032x" % (0xABC000 + i), len(d), partition_id="0") for i, d in enumerate(datas)]
    a, b = socket.socketpair()
    recv_stage = ChunkStage(0, max_batch_bytes=32 << 20, max_chunks=16, n_slots=1)
    if encrypt:
        recv_stage.set_e2ee_key(KEY)
    got = []

    def receiver():
        slot = recv_stage._free[-1]
        off = 0
        for _ in chunks:
            h, n = wire.recv_chunk(b, slot.out.view[off:])  # straight into page-locked memory
            got.append((h, off, n))
            off += native.round16(n)

    t = threading.Thread(target=receiver)
    t.start()
    try:
        slot = stage.begin()
        for d in datas:
            stage.add_bytes(slot, d)
        stage.launch(slot, compress=True, encrypt=encrypt)
        results = stage.collect(slot)  # views into the pinned output slot
        sent = wire.send_results(a, chunks, results)
        t.join(60)
        assert not t.is_alive() and sent == sum(53 + r.comp_len for r in results)
        rslot = recv_stage._free[-1]
        frames = [bytes(rslot.out.view[off: off + n]) for _, off, n in got]
        out = recv_stage.decode(frames, [h.raw_data_len for h, _, _ in got], encrypted=encrypt)
        for (h, _, n), c, d, r, (data, dg, st) in zip(got, chunks, datas, results, out):
            assert h.chunk_id == c.chunk_id and h.data_len == n == r.comp_len and h.raw_data_len == len(d) and h.is_compressed
            assert st == 0 and data == d and dg == r.md5 == hashlib.md5(d).digest()
    finally:
        a.close()
        b.close()
        recv_stage.close()
