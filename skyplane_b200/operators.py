"""Gateway operator plugin surface + the B200 compress/hash operator.

``GatewayOperator`` keeps the reference's contract (skyplane/gateway/operators/gateway_operator.py:32-122):
same constructor arguments, ``start_workers`` forks ``n_processes`` workers running
``worker_loop(worker_id, *self.args)``, ``process(chunk_req, *args) -> bool`` where True forwards
the request to ``output_queue``, False re-queues it, and an exception stops the gateway through
``error_event`` / ``error_queue``.

``GatewayCompressHash`` is the new stage (integration choice (ii) of SURVEY.md section 8b: a new
op_type placed between ``read_object_store`` and ``send``).  Per chunk it
  * reads ``<chunk_dir>/<chunk_id>.chunk`` into pinned memory (gateway_operator.py:350-352),
  * runs the fused LZ4-frame + MD5 kernel on the worker's GPU (worker_id -> device),
  * writes the frame to ``<chunk_id>.chunk.lz4`` and sets ``chunk.md5_hash`` (chunk.py:21),
  * reports ``compressed_size_bytes`` / ``uncompressed_size_bytes`` in the ``complete`` status record so the
    daemon's ``/api/v1/profile/compression`` endpoint lights up (gateway_daemon_api.py:130-134).
Its ``worker_loop`` drains a *batch* of requests per launch (the reference's loop sleeps 0.1 s per chunk,
gateway_operator.py:102, which would cap a worker below 10 chunks/s) but keeps the same state logging
and error conventions.  CUDA is initialised lazily inside the worker process, never in ``__init__``
(workers are forked, gateway_operator.py:66-70).
"""
from __future__ import annotations

import os
import queue
import time
import traceback
from abc import ABC, abstractmethod
from multiprocessing import Event, Process, Queue
from typing import List, Optional

from skyplane_b200.chunk import ChunkRequest, ChunkState
from skyplane_b200.chunk_store import ChunkStore
from skyplane_b200.gateway_queue import GatewayQueue


class GatewayOperator(ABC):
    def __init__(
        self,
        handle: str,
        region: str,
        input_queue: GatewayQueue,
        output_queue: Optional[GatewayQueue],
        error_event,
        error_queue: Queue,
        chunk_store: ChunkStore,
        n_processes: Optional[int] = 1,
    ):
        self.handle = handle
        self.region = region
        self.input_queue = input_queue
        self.output_queue = output_queue
        self.chunk_store = chunk_store
        self.error_event = error_event
        self.error_queue = error_queue
        self.n_processes = n_processes
        self.args = ()  # extra positional args handed to worker_loop / process
        self.processes: List[Process] = []
        self.exit_flags = [Event() for _ in range(self.n_processes)]
        self.worker_id: Optional[int] = None  # process-local

    def start_workers(self):
        for i in range(self.n_processes):
            p = Process(target=self.worker_loop, args=(i,) + self.args)
            p.start()
            self.processes.append(p)

    def stop_workers(self):
        for flag in self.exit_flags:
            flag.set()
        for p in self.processes:
            p.join()
        self.processes = []

    def _running(self, worker_id: int) -> bool:
        return not self.exit_flags[worker_id].is_set() and not self.error_event.is_set()

    def _fail(self, worker_id: int, exc: BaseException):
        print(f"[{self.handle}:{worker_id}] Exception: {exc}")
        self.error_queue.put(traceback.format_exc())
        self.error_event.set()
        self.exit_flags[worker_id].set()

    def worker_loop(self, worker_id: int, *args):
        """One request at a time, as the reference does (gateway_operator.py:79-115)."""
        self.worker_id = worker_id
        while self._running(worker_id):
            try:
                try:
                    chunk_req = self.input_queue.get_nowait(self.handle)
                except queue.Empty:
                    time.sleep(0.001)
                    continue
                self.chunk_store.log_chunk_state(chunk_req, ChunkState.in_progress, operator_handle=self.handle, worker_id=worker_id)
                if self.process(chunk_req, *args):
                    self.chunk_store.log_chunk_state(chunk_req, ChunkState.complete, operator_handle=self.handle, worker_id=worker_id)
                    if self.output_queue is not None:
                        self.output_queue.put(chunk_req)
                else:
                    time.sleep(0.1)
                    self.input_queue.put(chunk_req)
            except Exception as e:
                self._fail(worker_id, e)
        self.worker_exit(worker_id)

    def worker_exit(self, worker_id: int):
        pass

    @abstractmethod
    def process(self, chunk_req: ChunkRequest, *args) -> bool:
        ...


class GatewayCompressHash(GatewayOperator):
    """B200 stage: LZ4 frame + MD5 per chunk, batched per kernel launch."""

    def __init__(
        self,
        handle: str,
        region: str,
        input_queue: GatewayQueue,
        output_queue: Optional[GatewayQueue],
        error_event,
        error_queue: Queue,
        chunk_store: ChunkStore,
        n_processes: Optional[int] = 1,
        use_compression: Optional[bool] = True,
        max_batch_chunks: int = 64,
        max_batch_bytes: int = 512 << 20,
        n_gpus: Optional[int] = None,
        keep_frames_on_disk: bool = True,
        read_threads: int = 8,
        ingest_read_local: bool = True,
    ):
        super().__init__(handle, region, input_queue, output_queue, error_event, error_queue, chunk_store, n_processes)
        self.use_compression = use_compression
        self.max_batch_chunks = max_batch_chunks
        self.max_batch_bytes = max_batch_bytes
        self.n_gpus = n_gpus
        self.keep_frames_on_disk = keep_frames_on_disk
        self.read_threads = read_threads
        self.ingest_read_local = ingest_read_local
        self._stage = None  # created lazily in the worker process (fork + CUDA)
        self._readers = None  # thread pool for chunk-file reads, also per process

    # -- per-process GPU state ---------------------------------------------------------------
    def _get_stage(self):
        if self._stage is None:
            from skyplane_b200 import native
            from skyplane_b200.stage import ChunkStage

            ngpu = self.n_gpus or native.device_count()
            if ngpu <= 0:
                raise native.SkyChunkError(native.SKY_E_NOGPU, "GatewayCompressHash needs a CUDA device; there is no CPU fallback")
            device = (self.worker_id or 0) % ngpu
            if ngpu > 1:
                from skyplane_b200.numa import bind_to_gpu

                bind_to_gpu(device)  # pinned staging buffers on the GPU's own socket
            self._stage = ChunkStage(device, self.max_batch_bytes, self.max_batch_chunks, n_slots=2)
        return self._stage

    def worker_exit(self, worker_id: int):
        if self._readers is not None:
            self._readers.shutdown(wait=True)
            self._readers = None
        if self._stage is not None:
            self._stage.close()
            self._stage = None

    # -- the plugin method ---------------------------------------------------------------------
    def process(self, chunk_req: ChunkRequest, *args) -> bool:
        return self.process_batch([chunk_req])[0]

    # -- batch plumbing -------------------------------------------------------------------------
    def _read_into(self, path, view, n: int, offset: Optional[int] = None) -> bool:
        """Read exactly n bytes into pinned memory.  offset None: `path` is a chunk file that must hold exactly n
        bytes (False if it is not complete yet); otherwise the byte range [offset, offset+n) of a source object."""
        try:
            with open(path, "rb", buffering=0) as f:
                if offset:
                    f.seek(offset)
                got = 0
                while got < n:
                    r = f.readinto(view[got:])
                    if not r:
                        return False
                    got += r
                return True if offset is not None else not f.read(1)
        except FileNotFoundError:
            return False

    def _source_of(self, chunk_req: ChunkRequest):
        """-> (path, offset or None, ready).  ``src_type == "read_local"`` (skyplane/chunk.py:54) ingests the byte range
        straight from the source file into the staging slot -- no tmpfs chunk file in between (SURVEY.md section 8f
        row 3 for the POSIX case); every other request reads ``<chunk_id>.chunk`` as GatewaySender does."""
        chunk = chunk_req.chunk
        n = chunk.chunk_length_bytes
        if self.ingest_read_local and chunk_req.src_type == "read_local":
            off = chunk.file_offset_bytes or 0
            try:
                return chunk.src_key, off, os.stat(chunk.src_key).st_size >= off + n
            except FileNotFoundError:
                return chunk.src_key, off, False
        path = self.chunk_store.get_chunk_file_path(chunk.chunk_id)
        try:
            return path, None, os.stat(path).st_size == n  # upstream writes the file before queueing (gateway_operator.py:567-570)
        except FileNotFoundError:
            return path, None, False

    def _launch(self, reqs: List[ChunkRequest]):
        """Stage as many of `reqs` as fit one slot and launch them.
        -> (slot or None, launched indices, not-ready indices, leftover indices)"""
        stage = self._get_stage()
        slot = stage.begin()
        launched, not_ready, leftover, jobs = [], [], [], []
        for i, r in enumerate(reqs):
            chunk = r.chunk
            n = chunk.chunk_length_bytes
            if n > stage.max_batch_bytes:
                stage.release(slot)
                raise ValueError(f"chunk {chunk.chunk_id} ({n} B) exceeds the stage's max_batch_bytes")
            path, offset, ready = self._source_of(r)
            if not ready:
                not_ready.append(i)
            elif not stage.fits(slot, n):
                leftover.append(i)
            else:
                jobs.append((i, path, slot.reserve(n), n, offset))
        if jobs:
            # file -> pinned memory copies release the GIL: read the batch with a few threads
            if self._readers is None:
                from concurrent.futures import ThreadPoolExecutor

                self._readers = ThreadPoolExecutor(max_workers=self.read_threads)
            oks = list(self._readers.map(lambda j: self._read_into(j[1], j[2], j[3], j[4]), jobs))
            if not all(oks):  # a file changed under us: retry the whole batch later rather than hash partial data
                stage.release(slot)
                return None, [], [j[0] for j in jobs] + not_ready, leftover
            launched = [j[0] for j in jobs]
            stage.launch(slot)
            return slot, launched, not_ready, leftover
        stage.release(slot)
        if leftover:
            raise RuntimeError("staging slot cannot hold a single chunk")
        return None, [], not_ready, leftover

    def _finish(self, slot, reqs: List[ChunkRequest]):
        """Collect a launched batch: sets md5_hash, writes frames, attaches the size metadata."""
        results = self._get_stage().collect(slot)
        for r, res in zip(reqs, results):
            chunk = r.chunk
            chunk.md5_hash = res.md5
            if self.keep_frames_on_disk:
                with open(self.chunk_store.get_compressed_file_path(chunk.chunk_id), "wb") as f:
                    f.write(res.frame)
            r._stage_meta = {"compressed_size_bytes": res.comp_len, "uncompressed_size_bytes": res.raw_len}

    def process_batch(self, reqs: List[ChunkRequest]) -> List[bool]:
        """Compress + hash a batch synchronously. One bool per request (False = chunk file not ready yet, retry)."""
        ok = [True] * len(reqs)
        todo = list(range(len(reqs)))
        while todo:
            sub = [reqs[i] for i in todo]
            slot, launched, not_ready, leftover = self._launch(sub)
            for k in not_ready:
                ok[todo[k]] = False
            if slot is not None:
                self._finish(slot, [sub[k] for k in launched])
            todo = [todo[k] for k in leftover]
        return ok

    def _complete(self, worker_id: int, r: ChunkRequest):
        meta = r.__dict__.pop("_stage_meta", None)
        self.chunk_store.log_chunk_state(r, ChunkState.complete, operator_handle=self.handle, worker_id=worker_id, metadata=meta)
        if self.output_queue is not None:
            self.output_queue.put(r)

    def worker_loop(self, worker_id: int, *args):
        """Batch-draining, double-buffered loop with the reference's logging / error conventions: while the GPU
        works on one batch the next one is read from the chunk files into the other staging slot."""
        self.worker_id = worker_id
        inflight = []  # [(slot, reqs)] oldest first
        backlog: List[ChunkRequest] = []  # dequeued but not yet launched (did not fit the slot)
        try:
            while self._running(worker_id):
                try:
                    stage_free = self._stage is None or bool(self._stage._free)
                    if stage_free:
                        room = self.max_batch_chunks - len(backlog)
                        fresh = self.input_queue.get_batch_nowait(room, self.handle) if room > 0 else []
                        for r in fresh:
                            self.chunk_store.log_chunk_state(r, ChunkState.in_progress, operator_handle=self.handle, worker_id=worker_id)
                        cand = backlog + fresh
                        if cand:
                            slot, launched, not_ready, leftover = self._launch(cand)
                            if slot is not None:
                                inflight.append((slot, [cand[k] for k in launched]))
                            backlog = [cand[k] for k in leftover]
                            if not_ready:
                                time.sleep(0.1 if slot is None and not inflight else 0)
                                for k in not_ready:
                                    self.input_queue.put(cand[k])
                            if slot is not None and len(inflight) < 2:
                                continue  # try to get a second batch going before blocking on the first
                    if inflight:
                        slot, reqs = inflight.pop(0)
                        self._finish(slot, reqs)
                        for r in reqs:
                            self._complete(worker_id, r)
                    elif not backlog:
                        time.sleep(0.0005)
                except Exception as e:
                    self._fail(worker_id, e)
            # drain what is already on the GPU so no accepted chunk is lost on a clean stop
            if not self.error_event.is_set():
                for slot, reqs in inflight:
                    self._finish(slot, reqs)
                    for r in reqs:
                        self._complete(worker_id, r)
        finally:
            self.worker_exit(worker_id)


class ChecksumMismatchException(Exception):
    """Same name as skyplane/exceptions.py:44-48: the decoded chunk's MD5 differs from Chunk.md5_hash."""


class GatewayDecompressVerify(GatewayOperator):
    """Receiving side: what gateway_receiver.py:191-233 does after the socket read.

    ``<chunk_id>.chunk.lz4`` (the wire payload) -> LZ4 frame decode on the GPU -> ``<chunk_id>.chunk`` of exactly
    ``chunk_length_bytes`` bytes (the size check at gateway_receiver.py:213-218), and -- closing the reference's
    "# todo check hash" (gateway_receiver.py:231) -- the digest of the decoded bytes is compared with
    ``chunk.md5_hash`` when the sender supplied one.  A corrupt frame or a digest mismatch raises, which stops the
    gateway through ``error_event`` like any other operator failure."""

    def __init__(self, *args, max_batch_chunks: int = 64, max_batch_bytes: int = 512 << 20, n_gpus: Optional[int] = None,
                 remove_frames: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self.max_batch_chunks = max_batch_chunks
        self.max_batch_bytes = max_batch_bytes
        self.n_gpus = n_gpus
        self.remove_frames = remove_frames
        self._stage = None

    def _get_stage(self):
        if self._stage is None:
            from skyplane_b200 import native
            from skyplane_b200.stage import ChunkStage

            ngpu = self.n_gpus or native.device_count()
            if ngpu <= 0:
                raise native.SkyChunkError(native.SKY_E_NOGPU, "GatewayDecompressVerify needs a CUDA device; there is no CPU fallback")
            self._stage = ChunkStage((self.worker_id or 0) % ngpu, self.max_batch_bytes, self.max_batch_chunks, n_slots=1)
        return self._stage

    def worker_exit(self, worker_id: int):
        if self._stage is not None:
            self._stage.close()
            self._stage = None

    def process(self, chunk_req: ChunkRequest, *args) -> bool:
        from skyplane_b200 import native

        chunk = chunk_req.chunk
        fpath = self.chunk_store.get_compressed_file_path(chunk.chunk_id)
        if not fpath.exists():
            return False  # payload not received yet: retry (GatewayWaitReceiver semantics, gateway_operator.py:131-150)
        frame = fpath.read_bytes()
        (data, digest, status), = self._get_stage().decode([frame], [chunk.chunk_length_bytes])
        if status != 0:
            raise ValueError(f"chunk {chunk.chunk_id}: LZ4 frame rejected ({native.D_NAMES.get(status, status)})")
        if chunk.md5_hash is not None and bytes(chunk.md5_hash) != digest:
            raise ChecksumMismatchException(f"chunk {chunk.chunk_id}: md5 {digest.hex()} != expected {bytes(chunk.md5_hash).hex()}")
        with open(self.chunk_store.get_chunk_file_path(chunk.chunk_id), "wb") as f:
            f.write(data)
        if chunk.md5_hash is None:
            chunk.md5_hash = digest  # lets the upload step send Content-MD5 (gateway_operator.py:640)
        if self.remove_frames:
            fpath.unlink(missing_ok=True)
        return True
