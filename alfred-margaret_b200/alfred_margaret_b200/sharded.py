"""Sharded scans (multi-GPU / multi-window): the N > 1 host logic of SURVEY.md section 8e.

The haystack is cut into contiguous shards (`am_shard_plan`); rank r scans its shard plus a halo of
`halo_bytes` before it and reports only the matches that END inside its shard, so the per-rank lists
concatenate to exactly the single-shard list.  The only exchange is an all-gather of one match count per
rank (-> each rank's offset into the global list, and the total).  On GPUs that exchange is part of the C ABI
(`am_comm_init` + `am_find_all_sharded` / `am_count_sharded`: NCCL inside the library, queued on the scan's
stream; class `Comm` below binds it); the pure-host form (`exchange_counts` over any `torch.distributed`
backend) is what the CPU tests run over gloo.  Batches of independent haystacks need no collective at all.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Tuple

import numpy as np

from . import _ffi


def shard_plan(text_len: int, halo_bytes: int, n_shards: int, r: int) -> Tuple[int, int, int]:
    """(warm_begin, begin, end) of shard r: resident from warm_begin, reports matches ending in (begin, end]."""
    w, b, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _ffi.check(_ffi.lib().am_shard_plan(text_len, halo_bytes, n_shards, r, C.byref(w), C.byref(b), C.byref(e)))
    return w.value, b.value, e.value


def offsets_from_counts(counts, rank: int) -> Tuple[int, int]:
    """(this rank's offset into the global match list, total) from the gathered per-rank counts."""
    counts = [int(c) for c in counts]
    return sum(counts[:rank]), sum(counts)


def exchange_counts(n_local: int, rank: int, world: int, dist=None, device=None) -> Tuple[int, int, list]:
    """Host form of the exchange over torch.distributed (gloo in the CPU tests); returns (offset, total, counts)."""
    if world == 1 or dist is None:
        return 0, n_local, [n_local]
    import torch
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    allc = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine)
    counts = allc.tolist()
    off, total = offsets_from_counts(counts, rank)
    return off, total, counts


def find_all_sharded(scan: Callable[[int, int, int, int], np.ndarray], text_len: int, halo_bytes: int, rank: int, world: int,
                     dist=None, device=None):
    """Scan shard `rank` of a `text_len`-byte haystack.

    `scan(resident_begin, resident_end, report_begin, pos_base)` must return the ordered matches of the
    resident window [resident_begin, resident_end) whose end lies beyond `report_begin` (relative to the
    window), with positions rebased by `pos_base` -- i.e. am_find_all_dev on that window.
    Returns (matches, global_offset, total_matches).
    """
    w, b, e = shard_plan(text_len, halo_bytes, world, rank)
    matches = scan(w, e, b - w, w)
    off, total, _ = exchange_counts(len(matches), rank, world, dist, device)
    return matches, off, total


class Comm:
    """`am_comm`: one rank's end of the library's own exchange (NCCL, loaded by the library at run time)."""

    def __init__(self, rank: int, world: int, unique_id: bytes | None, device: int = -1):
        self.rank, self.world = rank, world
        h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, _ffi.COMM_ID_BYTES) if unique_id is not None else None
        _ffi.check(_ffi.lib().am_comm_init(rank, world, idbuf, device, C.byref(h)))
        self.handle = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_ffi.COMM_ID_BYTES)
        _ffi.check(_ffi.lib().am_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch(cls, dist, device: int):
        """Bootstrap over an initialised torch.distributed group: rank 0's id travels by broadcast_object_list."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 and world > 1 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        return cls(rank, world, box[0], device)

    def close(self):
        if self.handle:
            _ffi.lib().am_comm_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count(self, machine, dev_ptr, text_len, report_begin=0, pos_base=0, stream=None, case=None):
        t = _ffi.DevText(dev_ptr, text_len, report_begin, pos_base)
        r = _ffi.ShardResult()
        _ffi.check(_ffi.lib().am_count_sharded(machine.handle, machine._cs(case), self.handle, C.byref(t), stream, C.byref(r)))
        return r.n_local, r.global_offset, r.total

    def find_all(self, machine, dev_ptr, text_len, out_dev_ptr, cap, report_begin=0, pos_base=0, stream=None, case=None):
        t = _ffi.DevText(dev_ptr, text_len, report_begin, pos_base)
        r = _ffi.ShardResult()
        rc = _ffi.lib().am_find_all_sharded(machine.handle, machine._cs(case), self.handle, C.byref(t), stream, out_dev_ptr, cap, C.byref(r))
        if rc == _ffi.AM_E_OVERFLOW:
            raise OverflowError(r.n_local)
        _ffi.check(rc)
        return r.n_local, r.global_offset, r.total

    def contains_any(self, machine, dev_ptr, text_len, report_begin=0, pos_base=0, stream=None, case=None) -> bool:
        t = _ffi.DevText(dev_ptr, text_len, report_begin, pos_base)
        b = C.c_int()
        _ffi.check(_ffi.lib().am_contains_any_sharded(machine.handle, machine._cs(case), self.handle, C.byref(t), stream, C.byref(b)))
        return bool(b.value)

    def halo_exchange(self, dev_buf, halo_bytes, shard_len, stream=None):
        _ffi.check(_ffi.lib().am_shard_halo_exchange(self.handle, dev_buf, halo_bytes, shard_len, stream))

    def allreduce(self, value: int, op: str = "sum", stream=None) -> int:
        v = C.c_uint64(value)
        _ffi.check(_ffi.lib().am_comm_allreduce_u64(self.handle, C.byref(v), {"sum": 0, "max": 1, "min": 2}[op], stream))
        return v.value
