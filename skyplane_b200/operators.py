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
    ):
        super().__init__(handle, region, input_queue, output_queue, error_event, error_queue, chunk_store, n_processes)
        self.use_compression = use_compression
        self.max_batch_chunks = max_batch_chunks
        self.max_batch_bytes = max_batch_bytes
        self.n_gpus = n_gpus
        self.keep_frames_on_disk = keep_frames_on_disk
        self._stage = None  # created lazily in the worker process (fork + CUDA)

    # -- per-process GPU state ---------------------------------------------------------------
    def _get_stage(self):
        if self._stage is None:
            from skyplane_b200 import native
            from skyplane_b200.stage import ChunkStage

            ngpu = self.n_gpus or native.device_count()
            if ngpu <= 0:
                raise native.SkyChunkError(native.SKY_E_NOGPU, "GatewayCompressHash needs a CUDA device; there is no CPU fallback")
            device = (self.worker_id or 0) % ngpu
            self._stage = ChunkStage(device, self.max_batch_bytes, self.max_batch_chunks, n_slots=2)
        return self._stage

    def worker_exit(self, worker_id: int):
        if self._stage is not None:
            self._stage.close()
            self._stage = None

    # -- the plugin method ---------------------------------------------------------------------
    def process(self, chunk_req: ChunkRequest, *args) -> bool:
        return self.process_batch([chunk_req])[0]

    def process_batch(self, reqs: List[ChunkRequest]) -> List[bool]:
        """Compress + hash a batch. Returns one bool per request (False = chunk file not ready yet, retry)."""
        stage = self._get_stage()
        ok = [True] * len(reqs)
        pending = list(range(len(reqs)))
        while pending:
            slot = stage.begin()
            batch = []
            rest = []
            for i in pending:
                chunk = reqs[i].chunk
                path = self.chunk_store.get_chunk_file_path(chunk.chunk_id)
                n = chunk.chunk_length_bytes
                if n > stage.max_batch_bytes:
                    raise ValueError(f"chunk {chunk.chunk_id} ({n} B) exceeds the stage's max_batch_bytes")
                if not path.exists() or os.path.getsize(path) != n:
                    ok[i] = False  # upstream has not finished writing it (gateway_operator.py:131-150 semantics)
                    continue
                if not stage.fits(slot, n):
                    rest.append(i)
                    continue
                stage.add_file(slot, path, n)
                batch.append(i)
            if not batch:
                stage._free.append(slot)
                if rest:
                    raise RuntimeError("staging slot cannot hold a single chunk")
                break
            stage.launch(slot)
            results = stage.collect(slot)
            for i, r in zip(batch, results):
                chunk = reqs[i].chunk
                chunk.md5_hash = r.md5
                if self.keep_frames_on_disk:
                    with open(self.chunk_store.get_compressed_file_path(chunk.chunk_id), "wb") as f:
                        f.write(r.frame)
                reqs[i]._stage_meta = {"compressed_size_bytes": r.comp_len, "uncompressed_size_bytes": r.raw_len}
            pending = rest
        return ok

    def worker_loop(self, worker_id: int, *args):
        """Batch-draining loop with the reference's logging / error conventions."""
        self.worker_id = worker_id
        while self._running(worker_id):
            try:
                reqs = self.input_queue.get_batch_nowait(self.max_batch_chunks, self.handle)
                if not reqs:
                    time.sleep(0.0005)
                    continue
                for r in reqs:
                    self.chunk_store.log_chunk_state(r, ChunkState.in_progress, operator_handle=self.handle, worker_id=worker_id)
                oks = self.process_batch(reqs)
                retry = []
                for r, ok in zip(reqs, oks):
                    if ok:
                        meta = r.__dict__.pop("_stage_meta", None)
                        self.chunk_store.log_chunk_state(r, ChunkState.complete, operator_handle=self.handle, worker_id=worker_id, metadata=meta)
                        if self.output_queue is not None:
                            self.output_queue.put(r)
                    else:
                        retry.append(r)
                if retry:
                    time.sleep(0.1)
                    for r in retry:
                        self.input_queue.put(r)
            except Exception as e:
                self._fail(worker_id, e)
        self.worker_exit(worker_id)
