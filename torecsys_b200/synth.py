"""Deterministic, platform-independent synthetic data (counter-based hash, numpy only).

Used by tests, the golden-fixture generator and the small bench configurations so that inputs and
parameters can be re-created bit-for-bit anywhere instead of being stored.  Large tables in bench.py
are generated on the device with torch's generator instead (documented there).
"""
import zlib

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _seed_of(tag) -> np.uint64:
    if isinstance(tag, str):
        tag = zlib.crc32(tag.encode())
    return np.uint64(int(tag) & 0xFFFFFFFFFFFFFFFF)


def _hash64(count: int, tag) -> np.ndarray:
    """splitmix64 finaliser over the counters 0..count-1, keyed by `tag`."""
    with np.errstate(over='ignore'):
        z = np.arange(count, dtype=np.uint64) * _GOLD + _seed_of(tag) * _M2 + _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(shape, tag, lo=-1.0, hi=1.0, dtype=np.float32) -> np.ndarray:
    """Uniform [lo, hi) values with 24 random mantissa bits (exact in fp32), row-major order."""
    n = int(np.prod(shape)) if len(shape) else 1
    u = (_hash64(n, tag) >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(dtype).reshape(shape)


def integers(shape, tag, high, dtype=np.int64) -> np.ndarray:
    """Uniform integers in [0, high); `high` may be a scalar or broadcastable to `shape`."""
    n = int(np.prod(shape)) if len(shape) else 1
    z = (_hash64(n, tag) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    hi = np.broadcast_to(np.asarray(high, dtype=np.float64), shape).reshape(-1)
    out = np.minimum(np.floor(z * hi), hi - 1)
    return out.astype(dtype).reshape(shape)
