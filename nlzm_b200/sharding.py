"""Position-range sharding across the GPUs of one box: the input is replicated in
every GPU's HBM, rank r owns a contiguous range of positions and the engine builds its structures
over [begin - (W-1), end) itself, so there is no data-path collective; results are gathered to the
host by position."""
from __future__ import annotations

import numpy as np

ALIGN = 256   # RK256 block size; keeps shard boundaries on hash-block boundaries


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    per = -(-n // world)
    per = -(-per // ALIGN) * ALIGN
    b = min(rank * per, n)
    e = n if rank == world - 1 else min((rank + 1) * per, n)
    return b, e


def split_blocks(begin: int, end: int, max_block: int) -> list[tuple[int, int]]:
    """cut [begin, end) into engine calls of at most max_block positions"""
    out, b = [], begin
    while b < end:
        e = min(end, b + max_block)
        out.append((b, e))
        b = e
    return out or [(begin, end)]


def concat_views(parts):
    """[(begin, end, offsets, steps)] in any order -> whole-range CSR (offsets u64, dist u32, len u16)"""
    parts = sorted(parts, key=lambda p: p[0])
    offs, ds, ls, base = [np.zeros(1, np.uint64)], [], [], 0
    for (_, _, off, st) in parts:
        offs.append(off[1:].astype(np.uint64) + base)
        base += int(off[-1])
        ds.append(st["dist"])
        ls.append(st["len"])
    return np.concatenate(offs), np.concatenate(ds).astype(np.uint32), np.concatenate(ls).astype(np.uint16)
