"""Inter-operator queues of the gateway (API-compatible with skyplane/gateway/gateway_queue.py:4-61).

* ``GatewayQueue``    -- one bounded multiprocessing queue shared by the workers of the consuming operator.
* ``GatewayANDQueue`` -- fan-out: every registered consumer handle gets its own ``GatewayQueue`` and sees every item.

``get_batch_nowait`` is ours: the B200 operator drains many requests per kernel launch.
"""
from __future__ import annotations

import multiprocessing
import queue as _queue
from typing import Dict, Iterable, List

DEFAULT_DEPTH = 10000  # the reference's maxsize


class GatewayQueue:
    """FIFO of ChunkRequests between two operators."""

    def __init__(self, maxsize: int = DEFAULT_DEPTH):
        self.q = multiprocessing.Queue(maxsize)
        self.handles: List[str] = []

    # -- consumers ------------------------------------------------------------------------------
    def register_handle(self, requester_handle) -> None:
        self.handles.append(requester_handle)

    def get_handles(self) -> List[str]:
        return self.handles

    def get_nowait(self, requester_handle=None):
        """Next request or ``queue.Empty``."""
        return self.q.get_nowait()

    def get_batch_nowait(self, max_items: int, requester_handle=None) -> list:
        """Up to ``max_items`` requests that are available right now (possibly none)."""
        batch: list = []
        while len(batch) < max_items:
            try:
                batch.append(self.get_nowait(requester_handle))
            except _queue.Empty:
                break
        return batch

    def pop(self, requester_handle=None) -> None:
        self.q.get()

    # -- producers ------------------------------------------------------------------------------
    def put(self, chunk_req) -> None:
        self.q.put(chunk_req)

    def put_nowait(self, chunk_req) -> None:
        self.q.put_nowait(chunk_req)

    def put_many(self, chunk_reqs: Iterable) -> None:
        for r in chunk_reqs:
            self.q.put(r)

    def size(self) -> int:
        return self.q.qsize()


class GatewayANDQueue(GatewayQueue):
    """Broadcast queue: ``put`` delivers to every consumer's private queue (used behind ``mux_and``)."""

    def __init__(self, maxsize: int = DEFAULT_DEPTH):
        self.maxsize = maxsize
        self.q: Dict[str, GatewayQueue] = {}  # handle -> private queue
        self.temp_q = multiprocessing.Queue(maxsize)

    def register_handle(self, requester_handle) -> None:
        self.q[requester_handle] = GatewayQueue(self.maxsize)

    def get_handles(self) -> List[str]:
        return list(self.q)

    def get_handle_queue(self, requester_handle) -> GatewayQueue:
        return self.q[requester_handle]

    def get_nowait(self, requester_handle):
        return self.q[requester_handle].get_nowait()

    def pop(self, requester_handle) -> None:
        self.q[requester_handle].pop()

    def put(self, chunk_req) -> None:
        for private in self.q.values():
            private.put(chunk_req)

    def put_nowait(self, chunk_req) -> None:
        raise ValueError("GatewayANDQueue cannot be the first queue in a pipeline")

    def size(self) -> int:
        return max((private.size() for private in self.q.values()), default=0)
