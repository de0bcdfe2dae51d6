"""Wire framing straight from / into pinned staging memory.

Sender side replaces the tail of ``GatewaySender.process`` (skyplane/gateway/operators/gateway_operator.py:367-402):
``header = chunk.to_wire_header(...); header.to_socket(sock); sock.sendall(data)`` -- here ``data`` is a memoryview of
the stage's pinned output slot, so the frame is never copied into a Python ``bytes``.
Receiver side replaces the read loop of ``recv_chunks`` (skyplane/gateway/operators/gateway_receiver.py:150-189):
the payload is received directly into a caller-supplied (pinned) buffer in <= 4 MiB reads.
The 53-byte header layout is ``WireProtocolHeader``'s (skyplane/chunk.py:95-155).  With E2EE the payload is the sealed box
the stage produced (``data_len`` = box length, ``raw_data_len`` = original length, as gateway_operator.py:362-372 sets them).
"""
from __future__ import annotations

import socket
from typing import Optional, Sequence, Tuple

from skyplane_b200.chunk import Chunk, WireProtocolHeader

RECV_BLOCK = 4 << 20  # the reference's recv_block_size (gateway_receiver.py:40)


def send_chunk(sock: socket.socket, chunk: Chunk, payload, raw_len: int, n_chunks_left_on_socket: int = 0, is_compressed: bool = True) -> int:
    """Header + payload for one chunk. ``payload`` is any buffer (e.g. StageResult.frame). Returns bytes sent."""
    mv = memoryview(payload).cast("B")
    header = chunk.to_wire_header(n_chunks_left_on_socket=n_chunks_left_on_socket, wire_length=mv.nbytes, raw_wire_length=raw_len,
                                  is_compressed=is_compressed)
    header.to_socket(sock)
    sock.sendall(mv)
    return WireProtocolHeader.length_bytes() + mv.nbytes


def send_results(sock: socket.socket, chunks: Sequence[Chunk], results) -> int:
    """Pipeline a batch of StageResults onto one socket (n_chunks_left counts down like gateway_operator.py:368)."""
    total = 0
    n = len(chunks)
    for i, (c, r) in enumerate(zip(chunks, results)):
        total += send_chunk(sock, c, r.frame, r.raw_len, n_chunks_left_on_socket=n - i - 1, is_compressed=getattr(r, "is_compressed", True))
    return total


def recv_chunk(sock: socket.socket, into) -> Tuple[WireProtocolHeader, int]:
    """Read one header and its payload into ``into`` (a writable buffer of at least data_len bytes).
    Returns (header, data_len). Raises ConnectionError on EOF inside a chunk, ValueError if the buffer is too small."""
    header = WireProtocolHeader.from_socket(sock)
    mv = memoryview(into).cast("B")
    n = header.data_len
    if n > mv.nbytes:
        raise ValueError(f"payload of {n} bytes does not fit the {mv.nbytes}-byte receive buffer")
    got = 0
    while got < n:
        r = sock.recv_into(mv[got:n], min(n - got, RECV_BLOCK))
        if r == 0:
            raise ConnectionError("socket closed inside a chunk payload")
        got += r
    return header, n
