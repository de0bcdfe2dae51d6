"""Queues that connect gateway operators (mirror of skyplane/gateway/gateway_queue.py:4-61).

``GatewayQueue`` wraps one bounded ``multiprocessing.Queue``; ``GatewayANDQueue`` fans every item
out to one private queue per registered operator handle.  Kept API-compatible so the B200 operator
can sit in a stock ``gateway_daemon`` operator graph.  ``get_batch_nowait`` is an addition used by
the batch-draining worker loop.
"""
from __future__ import annotations

import queue
from multiprocessing import Queue
from typing import Dict, List


class GatewayQueue:
    def __init__(self, maxsize: int = 10000):
        self.q = Queue(maxsize)
        self.handles: List[str] = []

    def register_handle(self, requester_handle):
        self.handles.append(requester_handle)

    def get_handles(self):
        return self.handles

    def put(self, chunk_req):
        self.q.put(chunk_req)

    def put_nowait(self, chunk_req):
        self.q.put_nowait(chunk_req)

    def pop(self, requester_handle=None):
        self.q.get()

    def get_nowait(self, requester_handle=None):
        return self.q.get_nowait()  # raises queue.Empty

    def get_batch_nowait(self, max_items: int, requester_handle=None) -> list:
        """Drain up to ``max_items`` without blocking (may return [])."""
        out = []
        try:
            while len(out) < max_items:
                out.append(self.get_nowait(requester_handle))
        except queue.Empty:
            pass
        return out

    def size(self):
        return self.q.qsize()


class GatewayANDQueue(GatewayQueue):
    """Every downstream operator sees every chunk: one GatewayQueue per handle."""

    def __init__(self, maxsize: int = 10000):
        self.q: Dict[str, GatewayQueue] = {}
        self.maxsize = maxsize
        self.temp_q = Queue(maxsize)

    def register_handle(self, requester_handle):
        self.q[requester_handle] = GatewayQueue(self.maxsize)

    def get_handles(self):
        return list(self.q.keys())

    def get_handle_queue(self, requester_handle):
        return self.q[requester_handle]

    def put(self, chunk_req):
        for sub in self.q.values():
            sub.put(chunk_req)

    def put_nowait(self, chunk_req):
        raise ValueError("GatewayANDQueue cannot be the first queue in a pipeline")

    def pop(self, requester_handle):
        self.q[requester_handle].pop()

    def get_nowait(self, requester_handle):
        return self.q[requester_handle].get_nowait()

    def size(self):
        return max((sub.size() for sub in self.q.values()), default=0)
