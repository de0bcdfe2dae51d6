"""Chunk files + status log (mirror of skyplane/gateway/chunk_store.py:14-109).

A chunk's payload is the file ``<chunk_dir>/<chunk_id>.chunk`` (tmpfs in production,
compute/server.py:341).  Operators report state changes through ``log_chunk_state`` which feeds a
``multiprocessing.Queue`` drained by the gateway API.  The B200 stage additionally parks its output
frame next to the chunk as ``<chunk_id>.chunk.lz4`` (``get_compressed_file_path``).
"""
from __future__ import annotations

import os
import shutil
from datetime import datetime, timezone
from multiprocessing import Queue
from os import PathLike
from pathlib import Path
from typing import Dict, Optional

from skyplane_b200.chunk import ChunkRequest, ChunkState
from skyplane_b200.gateway_queue import GatewayQueue


class ChunkStore:
    def __init__(self, chunk_dir: PathLike):
        self.chunk_dir = Path(chunk_dir)
        self.chunk_dir.mkdir(parents=True, exist_ok=True)
        self.region_key_upload_id_mappings: Dict[str, str] = {}
        # a fresh store starts empty (chunk_store.py:21-24)
        for stale in list(self.chunk_dir.glob("*.chunk")) + list(self.chunk_dir.glob("*.chunk.lz4")):
            stale.unlink()
        self.chunk_requests: Dict[str, GatewayQueue] = {}  # partition -> queue of incoming requests
        self.chunk_status_queue: Queue = Queue()  # operator -> API status records

    def set_upload_ids_map(self, maps: Dict[str, str]):
        self.region_key_upload_id_mappings.update(maps)

    def get_upload_ids_map(self):
        return self.region_key_upload_id_mappings

    def add_partition(self, partition_id: str, queue: Optional[GatewayQueue] = None):
        if partition_id in self.chunk_requests:
            raise ValueError(f"Partition {partition_id} already exists")
        self.chunk_requests[partition_id] = queue if queue is not None else GatewayQueue()

    def add_chunk_request(self, chunk_request: ChunkRequest, state: ChunkState = ChunkState.registered):
        """Enqueue a request coming from the gateway API. Returns (queue size, accepted)."""
        part = chunk_request.chunk.partition_id
        if part not in self.chunk_requests:
            raise ValueError(f"Partition {part} does not exist in {self.chunk_requests} - was the gateway program loaded?")
        q = self.chunk_requests[part]
        try:
            q.put_nowait(chunk_request)
        except Exception as e:  # queue.Full
            print("Error adding chunk", e)
            return q.size(), False
        self.log_chunk_state(chunk_request, state)
        return q.size(), True

    def log_chunk_state(
        self,
        chunk_req: ChunkRequest,
        new_status: ChunkState,
        worker_id: Optional[int] = None,
        operator_handle: Optional[str] = None,
        metadata: Optional[Dict] = None,
    ):
        rec = {
            "chunk_id": chunk_req.chunk.chunk_id,
            "partition": chunk_req.chunk.partition_id,
            "state": new_status.name,
            "time": datetime.now(timezone.utc).replace(tzinfo=None).isoformat(),
            "handle": operator_handle,
            "worker_id": worker_id,
        }
        if metadata is not None:
            rec.update(metadata)
        self.chunk_status_queue.put(rec)

    def remaining_bytes(self) -> int:
        try:
            return shutil.disk_usage(self.chunk_dir).free
        except OSError:
            return 0

    def get_upload_id_map_path(self) -> Path:
        return self.chunk_dir / "upload_id_map.json"

    def get_chunk_file_path(self, chunk_id: str) -> Path:
        return self.chunk_dir / f"{chunk_id}.chunk"

    def get_compressed_file_path(self, chunk_id: str) -> Path:
        return self.chunk_dir / f"{chunk_id}.chunk.lz4"
