"""Deterministic synthetic chunk generators for the BASELINE.json configurations.

  random_chunk(i, n)        config 2/4: uniform random bytes, chunk i = default_rng(1000+i).bytes(n)
  silesia_like_chunk(i, n)  config 3:   seeded mix of text-like, record-like, numeric, constant and random
                            segments (4 KiB - 1 MiB pieces); liblz4 level-0 ratio lands around 2-3,
                            the neighbourhood of the Silesia corpus with ``lz4 -1``.
Pure numpy; used by tests and bench.py (data generation is always outside timed regions).
"""
from __future__ import annotations

import numpy as np

_WORDS_CACHE = {}


def random_chunk(i: int, n: int) -> bytes:
    return np.random.default_rng(1000 + i).bytes(n)


def _vocab(seed: int = 7, size: int = 4096):
    if seed not in _WORDS_CACHE:
        r = np.random.default_rng(seed)
        letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
        p = np.arange(26, 0, -1, dtype=np.float64)
        p /= p.sum()
        words = []
        for _ in range(size):
            L = int(r.integers(2, 11))
            words.append(bytes(r.choice(letters, size=L, p=p)))
        _WORDS_CACHE[seed] = words
    return _WORDS_CACHE[seed]


def _text(r: np.random.Generator, n: int) -> bytes:
    words = _vocab()
    idx = (r.zipf(1.25, size=n // 4 + 16) - 1) % len(words)
    seps = r.choice(np.frombuffer(b"     ,.\n", dtype=np.uint8), size=idx.size)
    parts = []
    total = 0
    for k, s in zip(idx, seps):
        w = words[k]
        parts.append(w)
        parts.append(bytes([s]))
        total += len(w) + 1
        if total >= n:
            break
    return b"".join(parts)[:n]


def _records(r: np.random.Generator, n: int) -> bytes:
    out = []
    total = 0
    k = int(r.integers(0, 1 << 20))
    while total < n:
        rec = b'{"id": %d, "ts": %d, "lat": %.4f, "lon": %.4f, "status": "%s", "tags": ["alpha", "beta"]}\n' % (
            k, 1600000000 + 17 * k, float(r.uniform(-90, 90)), float(r.uniform(-180, 180)), (b"ok", b"late", b"lost")[k % 3])
        out.append(rec)
        total += len(rec)
        k += 1
    return b"".join(out)[:n]


def _numeric(r: np.random.Generator, n: int) -> bytes:
    # sensor-like little-endian int32 samples: a slow random walk with plateaus
    m = n // 4 + 1
    steps = r.choice(np.array([0, 0, 0, 0, 1, -1, 2], dtype=np.int64), size=m)
    walk = np.cumsum(steps) + int(r.integers(0, 1 << 16))
    return walk.astype("<i4").tobytes()[:n]


def _constant(r: np.random.Generator, n: int) -> bytes:
    return bytes([int(r.integers(0, 256))]) * n


def _noise(r: np.random.Generator, n: int) -> bytes:
    return r.bytes(n)


_KINDS = (_text, _records, _numeric, _constant, _noise)
_WEIGHTS = np.array([0.42, 0.22, 0.18, 0.10, 0.08])  # text, records, numeric, constant, noise


def silesia_like_chunk(i: int, n: int) -> bytes:
    r = np.random.default_rng(2000 + i)
    out = []
    total = 0
    while total < n:
        kind = _KINDS[int(r.choice(len(_KINDS), p=_WEIGHTS))]
        hi = 65536 if kind is _constant else (1 << 20)
        seg = int(min(n - total, r.integers(4096 if kind is not _constant else 64, hi + 1)))
        out.append(kind(r, seg))
        total += seg
    data = b"".join(out)
    assert len(data) == n
    return data
