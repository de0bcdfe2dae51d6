"""GPU test of the plugin boundary: GatewayCompressHash workers inside the queue harness
(BASELINE config 1/4 shape, small): chunk files -> GatewayQueue -> forked worker -> GPU -> output queue."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import pytest

import oracle
import oracle.reflib as ref
from skyplane_b200 import synth
from skyplane_b200.chunk import WireProtocolHeader

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]
ROOT = Path(__file__).resolve().parent.parent

DRIVER = r"""
import json, sys
from pathlib import Path
from skyplane_b200.harness import run_stream
base = Path(sys.argv[1]); n_req = int(sys.argv[2]); workers = int(sys.argv[3])
files = sorted((base / "pool").glob("*.bin"), key=lambda p: int(p.stem))
lens = [p.stat().st_size for p in files]
res = run_stream(base / "chunks", files, lens, n_req, n_workers=workers, max_batch_chunks=8, max_batch_bytes=64 << 20, keep_frames=True)
print("RESULT " + json.dumps(res))
"""


def _run(base: Path, n_req: int, workers: int):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, "-c", DRIVER, str(base), str(n_req), str(workers)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_operator_in_queue_harness():
    base = Path(tempfile.mkdtemp(prefix="skyb200_test_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        (base / "pool").mkdir()
        pool = [synth.random_chunk(0, 8 << 20), synth.silesia_like_chunk(1, 8 << 20), b"", b"x" * 13, synth.silesia_like_chunk(2, (1 << 20) + 77),
                synth.random_chunk(5, 65536)]
        for k, d in enumerate(pool):
            (base / "pool" / f"{k}.bin").write_bytes(d)
        n_req = 40
        res = _run(base, n_req, workers=2)
        assert len(res["records"]) == n_req and res["bytes"] == sum(len(pool[i % len(pool)]) for i in range(n_req))
        assert res["status"].get("complete") == n_req and res["status"].get("in_progress") == n_req and res["status"].get("registered") == n_req
        assert res["uncompressed_bytes"] == res["bytes"] and 0 < res["compressed_bytes"] < res["bytes"]
        for rec in res["records"]:
            data = pool[rec["pool_index"]]
            assert rec["md5"] == hashlib.md5(data).hexdigest()
            frame = Path(rec["frame_path"]).read_bytes()
            assert oracle.lz4f_decode(frame, len(data)) == data
            assert ref.lz4f_decompress(frame, len(data)) == data  # what the destination gateway does (gateway_receiver.py:196)
            # the three header fields the sender derives (gateway_operator.py:367-372)
            hdr = WireProtocolHeader(rec["chunk_id"], len(frame), rec["raw_len"], True, 0)
            back = WireProtocolHeader.from_bytes(hdr.to_bytes())
            assert back.data_len == len(frame) and back.raw_data_len == len(data) and back.is_compressed
    finally:
        import shutil

        shutil.rmtree(base, ignore_errors=True)


def test_ingest_without_chunk_files():
    """SURVEY section 8f row 3 on the GPU: (a) `read_local` requests are byte ranges of one source object, read straight
    into the staging slot by the operator (no <chunk_id>.chunk); (b) a streaming body (readinto / read) fills a slot through
    ChunkStage.add_stream.  Digests equal hashlib's over exactly the requested range (s3_interface.py:178-194)."""
    import io
    import multiprocessing as mp

    from skyplane_b200.chunk import Chunk, ChunkRequest
    from skyplane_b200.chunk_store import ChunkStore
    from skyplane_b200.gateway_queue import GatewayQueue
    from skyplane_b200.operators import GatewayCompressHash
    from skyplane_b200.stage import ChunkStage

    base = Path(tempfile.mkdtemp(prefix="skyb200_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        obj = synth.silesia_like_chunk(7, 5 << 20) + synth.random_chunk(8, (3 << 20) + 1234)
        src = base / "object.bin"
        src.write_bytes(obj)
        cs = ChunkStore(base / "chunks")
        qin, qout = GatewayQueue(), GatewayQueue()
        err_ev, err_q = mp.Event(), mp.Queue()
        op = GatewayCompressHash("ch", "test:r", qin, qout, err_ev, err_q, cs, n_processes=1, max_batch_chunks=8, max_batch_bytes=32 << 20)
        ranges = [(0, 1 << 20), (1 << 20, (2 << 20) + 77), ((3 << 20) + 77, len(obj) - (3 << 20) - 77), (12345, 1), (len(obj) - 13, 13)]
        # (the plugin method is called in-process: this pytest process already owns a CUDA context, so it must not fork)
        op.worker_id = 0
        try:
            reqs = [ChunkRequest(Chunk(str(src), "dst", "%032x" % (0xFEED00 + i), n, partition_id="0", file_offset_bytes=off), src_type="read_local")
                    for i, (off, n) in enumerate(ranges)]
            assert op.process_batch(reqs) == [True] * len(reqs)
            assert op.process(ChunkRequest(Chunk(str(src), "dst", "%032x" % 0xFEEDFF, 10, partition_id="0", file_offset_bytes=len(obj) - 5),
                                           src_type="read_local")) is False  # range not (yet) inside the object: retry, not an error
            for r, (off, n) in zip(reqs, ranges):
                data = obj[off: off + n]
                assert r.chunk.md5_hash == hashlib.md5(data).digest()
                assert oracle.lz4f_decode(cs.get_compressed_file_path(r.chunk.chunk_id).read_bytes(), n) == data
                assert not cs.get_chunk_file_path(r.chunk.chunk_id).exists()  # nothing was staged on tmpfs
        finally:
            op.worker_exit(0)
        # (b) streaming bodies
        stage = ChunkStage(0, max_batch_bytes=32 << 20, max_chunks=8, n_slots=1)
        try:
            class ReadOnly:  # a body without readinto, like botocore's StreamingBody
                def __init__(self, b):
                    self.f = io.BytesIO(b)

                def read(self, n):
                    return self.f.read(min(n, 70000))

            slot = stage.begin()
            stage.add_stream(slot, io.BytesIO(obj[: 3 << 20]), 3 << 20)
            stage.add_stream(slot, ReadOnly(obj[3 << 20:]), len(obj) - (3 << 20))
            with pytest.raises(EOFError):
                stage.add_stream(slot, io.BytesIO(b"short"), 100)
            stage.launch(slot)
            a, b = stage.collect(slot)
            assert a.md5 == hashlib.md5(obj[: 3 << 20]).digest() and b.md5 == hashlib.md5(obj[3 << 20:]).digest()
            assert oracle.lz4f_decode(bytes(a.frame), 3 << 20) == obj[: 3 << 20]
            assert oracle.lz4f_decode(bytes(b.frame), len(obj) - (3 << 20)) == obj[3 << 20:]
        finally:
            stage.close()
    finally:
        import shutil

        shutil.rmtree(base, ignore_errors=True)
