"""Synthetic workloads of BASELINE.json's configs (bench / test tooling).

Host (numpy) twins of the device generators in csrc/am_synth.cu: every byte is a pure function of
(seed, absolute index), so any slice of a multi-GiB device-generated haystack can be reproduced
on the host and handed to the CPU oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
AZ = bytes(range(ord("a"), ord("z") + 1))


def _mix(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(33); x *= np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33); x *= np.uint64(0xC4CEB9FE1A85EC53)
    x ^= x >> np.uint64(33)
    return x


def fill_host(first: int, length: int, seed: int, alphabet: bytes = AZ) -> np.ndarray:
    """Bytes [first, first + length) of the synthetic text."""
    with np.errstate(over="ignore"):
        w0, w1 = first >> 3, (first + length + 7) >> 3
        words = np.arange(w0, w1, dtype=np.uint64)
        r = _mix(np.uint64(seed) ^ (words * GOLDEN))
        r8 = r.view(np.uint8).reshape(-1, 8).astype(np.uint32)  # little-endian byte j of each word
        alpha = np.frombuffer(alphabet, dtype=np.uint8)
        out = alpha[(r8 * np.uint32(len(alphabet))) >> 8].reshape(-1)
    lo = first - (w0 << 3)
    return np.ascontiguousarray(out[lo:lo + length])


def plant_host(buf: np.ndarray, first: int, seed: int, needles, block: int = 4096) -> None:
    """In-place twin of synth_plant_kernel on bytes [first, first + len(buf))."""
    n = len(needles)
    length = buf.size
    k0 = first // block
    kfirst = max(k0 - 1, 0)
    klast = (first + length) // block
    ks = np.arange(kfirst, klast + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        r = _mix(np.uint64(seed) ^ (ks * GOLDEN))
    ids = ((r & np.uint64(0xFFFFFFFF)) % np.uint64(n)).astype(np.int64)
    offs = (np.uint64(16) + (r >> np.uint64(32)) % np.uint64(block - 16)).astype(np.int64)
    for k, i, o in zip(ks.astype(np.int64).tolist(), ids.tolist(), offs.tolist()):
        nb = np.frombuffer(needles[i], dtype=np.uint8)
        at = k * block + o - first
        lo, hi = max(at, 0), min(at + nb.size, length)
        if lo < hi:
            buf[lo:hi] = nb[lo - at:hi - at]


def random_needles(n: int, seed: int, min_len: int = 4, max_len: int = 16, alphabet: bytes = AZ):
    """`n` distinct random needles, lengths uniform in [min_len, max_len] (config C2 of SURVEY.md 8d)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    out, seen = [], set()
    while len(out) < n:
        ln = int(rng.integers(min_len, max_len + 1))
        b = alpha[rng.integers(0, len(alphabet), size=ln)].tobytes()
        if b not in seen:
            seen.add(b)
            out.append(b)
    return out


def fill_dev(dev_ptr: int, length: int, first: int, seed: int, alphabet: bytes = AZ, stream=None) -> None:
    _ffi.check(_ffi.lib().am_synth_fill_dev(dev_ptr, length, first, seed, alphabet, len(alphabet), stream))


def plant_dev(dev_ptr: int, length: int, first: int, seed: int, needles, block: int = 4096, stream=None) -> None:
    keep = [C.create_string_buffer(b, len(b)) for b in needles]
    arr = (_ffi.U8Slice * len(needles))()
    for i, (b, k) in enumerate(zip(needles, keep)):
        arr[i] = _ffi.U8Slice(C.cast(k, C.c_void_p).value, 0, len(b))
    _ffi.check(_ffi.lib().am_synth_plant_dev(dev_ptr, length, first, seed, arr, len(needles), block, stream))
