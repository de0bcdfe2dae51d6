"""Chunk metadata types the gateway stage speaks.

Same names, fields, defaults and wire layout as the reference's ``skyplane/chunk.py`` (Chunk :9-43,
ChunkRequest :47-76, ChunkState :79-92, WireProtocolHeader :95-167) so that objects pickled through
a ``GatewayQueue`` or JSON-ed by the gateway API are interchangeable with the reference's.  The
53-byte header is packed with one ``struct`` format instead of byte-string concatenation.

One addition: ``Chunk.as_json_dict()`` / ``Chunk.from_json_dict()`` hex-encode ``md5_hash`` because
the reference JSON-dumps ``Chunk.as_dict()`` (gateway_operator.py:299) and raw ``bytes`` are not
JSON-serialisable -- the reason the reference leaves the digest unset (gateway_operator.py:577-582).
"""
from __future__ import annotations

import socket
import struct
from dataclasses import asdict, dataclass, fields
from enum import Enum, auto
from functools import total_ordering
from typing import Dict, Optional


@dataclass
class Chunk:
    """A contiguous piece of an object (an object is one or more chunks)."""

    src_key: str
    dest_key: str
    chunk_id: str  # 32 hex digits (uuid4().hex in the reference's Chunker)
    chunk_length_bytes: int
    partition_id: Optional[str] = None
    mime_type: Optional[str] = None

    md5_hash: Optional[bytes] = None  # 16 raw bytes; filled by the B200 stage

    multi_part: Optional[bool] = False
    file_offset_bytes: Optional[int] = None
    part_number: Optional[int] = None
    upload_id: Optional[str] = None

    def to_wire_header(self, n_chunks_left_on_socket: int, wire_length: int, raw_wire_length: int, is_compressed: bool = False):
        return WireProtocolHeader(
            chunk_id=self.chunk_id,
            data_len=wire_length,
            raw_data_len=raw_wire_length,
            is_compressed=is_compressed,
            n_chunks_left_on_socket=n_chunks_left_on_socket,
        )

    def as_dict(self) -> Dict:
        return asdict(self)

    @staticmethod
    def from_dict(d: Dict) -> "Chunk":
        return Chunk(**d)

    # -- JSON-safe variants (md5 as hex) ------------------------------------------------------
    def as_json_dict(self) -> Dict:
        d = asdict(self)
        if d["md5_hash"] is not None:
            d["md5_hash"] = bytes(d["md5_hash"]).hex()
        return d

    @staticmethod
    def from_json_dict(d: Dict) -> "Chunk":
        d = dict(d)
        if isinstance(d.get("md5_hash"), str):
            d["md5_hash"] = bytes.fromhex(d["md5_hash"])
        known = {f.name for f in fields(Chunk)}
        return Chunk(**{k: v for k, v in d.items() if k in known})


@dataclass
class ChunkRequest:
    """Gateway-local state wrapped around a Chunk; this is what ``process()`` receives."""

    chunk: Chunk
    src_region: Optional[str] = None
    dst_region: Optional[str] = None
    src_type: Optional[str] = None  # "object_store" | "random" | "read_local"
    dst_type: Optional[str] = None  # "object_store" | "save_local"
    src_random_size_mb: Optional[int] = None
    src_object_store_bucket: Optional[str] = None
    dst_object_store_bucket: Optional[str] = None

    def __post_init__(self):
        if self.src_type == "object_store":
            assert self.src_object_store_bucket is not None
        elif self.src_type == "random":
            assert self.src_random_size_mb is not None
        if self.dst_type == "object_store":
            assert self.dst_object_store_bucket is not None

    def as_dict(self) -> Dict:
        out = asdict(self)
        out["chunk"] = self.chunk.as_dict()
        return out

    @staticmethod
    def from_dict(in_dict: Dict) -> "ChunkRequest":
        # the gateway API posts bare Chunk dicts; the request wrapper is local (chunk.py:73-76)
        return ChunkRequest(chunk=Chunk.from_dict(in_dict))


@total_ordering
class ChunkState(Enum):
    registered = auto()
    in_progress = auto()
    failed = auto()
    queued = auto()
    complete = auto()

    @staticmethod
    def from_str(s: str) -> "ChunkState":
        return ChunkState[s.lower()]

    def __lt__(self, other):
        return self.value < other.value


_WIRE = struct.Struct(">QI16sQQBQ")  # magic, version, chunk_id, data_len, raw_data_len, is_compressed, n_left
_MAGIC = 0x534B595F4C41524B  # "SKY_LARK"
_VERSION = 3  # v3 = uuid chunk ids


@dataclass
class WireProtocolHeader:
    """53-byte big-endian header that precedes each chunk on a gateway-to-gateway socket."""

    chunk_id: str  # 128-bit id as 32 hex digits
    data_len: int  # bytes on the wire (compressed [+ encrypted])
    raw_data_len: int  # original chunk bytes
    is_compressed: bool
    n_chunks_left_on_socket: int

    @staticmethod
    def magic_hex() -> int:
        return _MAGIC

    @staticmethod
    def protocol_version() -> int:
        return _VERSION

    @staticmethod
    def length_bytes() -> int:
        return _WIRE.size

    @staticmethod
    def from_bytes(data: bytes) -> "WireProtocolHeader":
        assert len(data) == _WIRE.size, f"{len(data)} != {_WIRE.size}"
        magic, version, cid, data_len, raw_len, comp, n_left = _WIRE.unpack(data)
        if magic != _MAGIC:
            raise ValueError(f"Invalid magic number, got {magic:x} but expected {_MAGIC:x}")
        if version != _VERSION:
            raise ValueError(f"Invalid protocol version, got {version} but expected {_VERSION}")
        return WireProtocolHeader(
            chunk_id=cid.hex(), data_len=data_len, raw_data_len=raw_len, is_compressed=bool(comp), n_chunks_left_on_socket=n_left
        )

    def to_bytes(self) -> bytes:
        cid = bytes.fromhex(self.chunk_id)
        assert len(cid) == 16
        return _WIRE.pack(_MAGIC, _VERSION, cid, self.data_len, self.raw_data_len, int(bool(self.is_compressed)), self.n_chunks_left_on_socket)

    @staticmethod
    def from_socket(sock: socket.socket) -> "WireProtocolHeader":
        want = _WIRE.size
        buf = bytearray()
        while len(buf) < want:
            part = sock.recv(want - len(buf))
            if not part:
                raise ConnectionError("socket closed inside a chunk header")
            buf += part
        return WireProtocolHeader.from_bytes(bytes(buf))

    def to_socket(self, sock: socket.socket):
        assert sock.sendall(self.to_bytes()) is None
