"""Sharded scans (multi-GPU / multi-window): the N > 1 host logic of SURVEY.md section 8e.

The haystack is cut into contiguous shards (`am_shard_plan`); rank r scans its shard plus a halo of
`halo_bytes` before it and reports only the matches that END inside its shard, so the per-rank lists
concatenate to exactly the single-shard list.  The only exchange is an all-gather of one match count per
rank (-> each rank's offset into the global list, and the total): `torch.distributed` over NCCL on GPUs,
gloo in the CPU tests.  Batches of independent haystacks need no collective at all.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np

from . import _ffi


def shard_plan(text_len: int, halo_bytes: int, n_shards: int, r: int) -> Tuple[int, int, int]:
    """(warm_begin, begin, end) of shard r: resident from warm_begin, reports matches ending in (begin, end]."""
    w, b, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _ffi.check(_ffi.lib().am_shard_plan(text_len, halo_bytes, n_shards, r, C.byref(w), C.byref(b), C.byref(e)))
    return w.value, b.value, e.value


def exchange_counts(n_local: int, rank: int, world: int, dist=None, device=None) -> Tuple[int, int, list]:
    """All-gather the per-shard match counts; returns (this rank's offset in the global list, total, counts)."""
    if world == 1 or dist is None:
        return 0, n_local, [n_local]
    import torch
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    allc = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine)
    counts = allc.tolist()
    return int(sum(counts[:rank])), int(sum(counts)), counts


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
