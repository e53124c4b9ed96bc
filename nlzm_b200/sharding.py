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


# --------------------------------------------------------------------------------------------------
# Position sharding with the segment hand-over (include/nlzm_mf.h "segments"): every rank ranks and
# merges ITS OWN positions only; the window behind its range comes from the ranks that own it, as a
# one-sided copy of their sorted blocks + pointers (peer copy / CUDA IPC over NVLink, or staged through
# the host). No collective on the data path: the only group call exchanges a few hundred bytes of
# descriptors (and doubles as the "neighbours are ready" barrier).
# --------------------------------------------------------------------------------------------------

def blocks_for(begin: int, end: int, window: int, max_block: int = 1 << 28) -> list[tuple[int, int]]:
    """engine calls for [begin, end): blocks of at least one window (a later block then only needs the blocks
    of its own shard behind it) and at most 2^28 positions"""
    blk = min(max_block, max(window, 1 << 27))
    return split_blocks(begin, end, blk)


class ShardedFind:
    """Drives one engine (one rank) through   prepare(first block) -> later blocks -> exchange -> first block.

    transport: "ipc" (other processes on the same box), "peer" (engines of one process), "host" (staged copy)
    group:     torch.distributed group used for the descriptor exchange (None = single rank)
    """

    def __init__(self, mf, rank: int, world: int, window: int, group=None, transport: str = "ipc"):
        self.mf, self.rank, self.world, self.W, self.group, self.transport = mf, rank, world, window, group, transport
        self.imported_bytes = 0
        self._fast = transport != "host"  # tensor collectives (tests that replace _exchange switch this off)
        self.ms = {}                  # host wall time per protocol phase, accumulated over run() calls

    def _tick(self, name, t0):
        import time
        t1 = time.perf_counter()
        self.ms[name] = self.ms.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    def _exchange(self, payload):
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, payload, group=self.group)
        return out

    MAX_SEGS = 12

    def _exchange_descs(self, err, mine):
        """descriptors of every rank in ONE fixed-size all_gather (a few KB over the host-side group);
        falls back to the object exchange when there is an error text or a host copy to ship"""
        import ctypes
        import torch
        import torch.distributed as dist
        from ._lib import SegmentDesc
        self.DESC_BYTES = ctypes.sizeof(SegmentDesc)
        slow = err is not None or any("host" in m for m in mine) or len(mine) > self.MAX_SEGS
        rec = 8 + self.MAX_SEGS * self.DESC_BYTES
        buf = torch.zeros(rec, dtype=torch.uint8)
        buf[0] = 1 if slow else 0
        if not slow:
            buf[1] = len(mine)
            for i, m in enumerate(mine):
                assert len(m["desc"]) == self.DESC_BYTES
                buf[8 + i * self.DESC_BYTES: 8 + (i + 1) * self.DESC_BYTES] = torch.frombuffer(bytearray(m["desc"]), dtype=torch.uint8)
        allb = [torch.empty(rec, dtype=torch.uint8) for _ in range(self.world)]
        dist.all_gather(allb, buf, group=self.group)
        if any(int(b[0]) for b in allb):                 # somebody needs the general path: everybody takes it
            return self._exchange({"err": err, "segs": mine})
        out = []
        for b in allb:
            segs = []
            for i in range(int(b[1])):
                raw = bytes(b[8 + i * self.DESC_BYTES: 8 + (i + 1) * self.DESC_BYTES].numpy())
                d = SegmentDesc.from_buffer_copy(raw)
                segs.append({"desc": raw, "pos": (int(d.pos_begin), int(d.pos_end))})
            out.append({"err": None, "segs": segs})
        return out

    def _agree_fast(self, err):
        """one small all_reduce; the error texts are only exchanged when there is an error somewhere"""
        import torch
        import torch.distributed as dist
        flag = torch.tensor([1 if err else 0], dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()):
            self._agree(err)

    def _agree(self, err):
        """every rank learns whether any rank failed in the phase just finished (also a barrier)"""
        errs = [e for e in self._exchange(err) if e]
        if errs:
            raise RuntimeError("sharded find failed on some rank: " + "; ".join(errs))

    def run(self, blocks, find):
        """blocks: this rank's engine calls in position order; find(b, e, i) does the engine call and consumes
        its result. Returns the list of find() results in block order. A failure on any rank raises on every
        rank at the same point of the protocol (no rank is left waiting in a group call)."""
        mf = self.mf
        if self.world == 1:
            return [find(b, e, i) for i, (b, e) in enumerate(blocks)]
        import time
        first = blocks[0]
        need_from = max(0, first[0] - (self.W - 1))
        err, later, mine = None, [], []
        t = time.perf_counter()
        try:
            mf.prepare(*first)
            t = self._tick("prepare", t)
            later = [find(b, e, i + 1) for i, (b, e) in enumerate(blocks[1:])]
            t = self._tick("later_blocks", t)
            reach = max(0, blocks[-1][1] - (self.W - 1))              # the next shard reaches no further back
            if self.transport == "ipc":
                descs = mf.publish_segments(reach) if self.rank + 1 < self.world else []
            else:
                if self.rank + 1 < self.world:
                    mf.trim_segments(reach)
                descs = mf.export_segments()
            for d in descs:
                mine.append({"desc": bytes(d), "pos": (int(d.pos_begin), int(d.pos_end))})
            if self.transport == "host":
                # staged copy: every rank publishes the slices the next ranks can reach
                nxt = blocks[-1][1]
                for i, (d, item) in enumerate(zip(descs, mine)):
                    if item["pos"][1] > nxt - (self.W - 1) - 1:
                        item["host"] = mf.read_segment(i, d)
            t = self._tick("export", t)
        except Exception as ex:  # noqa: BLE001 - reported to every rank below
            err = f"rank {self.rank}: {ex}"
        everyone = self._exchange_descs(err, mine) if self._fast else self._exchange({"err": err, "segs": mine})
        t = self._tick("exchange_wait", t)
        errs = [p["err"] for p in everyone if p["err"]]
        if errs:
            raise RuntimeError("sharded find failed on some rank: " + "; ".join(errs))
        self.imported_bytes = 0
        try:
            for q in range(self.rank - 1, -1, -1):
                for item in everyone[q]["segs"]:
                    pb, pe = item["pos"]
                    if pe > need_from and pe <= first[0]:
                        if self.transport == "host":
                            mf.import_segment(item["desc"], host_copy=item["host"])
                        else:
                            # asynchronous: the copies overlap this rank's HT / RK stages (find waits for them)
                            mf.import_segment(item["desc"], via=(1 if self.transport == "ipc" else 0) | 0x100)
                        self.imported_bytes += (pe - pb) * 48
        except Exception as ex:  # noqa: BLE001
            err = f"rank {self.rank}: {ex}"
        t = self._tick("import", t)
        # every rank learns whether an import failed anywhere. (Asynchronous copies may still be in flight here: the
        # export buffers are only rewritten by the next publish, which comes after the agreement at the END of this
        # run, and every rank's find — which waits for its imports — has returned by then.)
        (self._agree_fast if self._fast else self._agree)(err)
        t = self._tick("agree_wait", t)
        try:
            res = find(first[0], first[1], 0)
        except Exception as ex:  # noqa: BLE001
            err = f"rank {self.rank}: {ex}"
            res = None
        t = self._tick("finish_first", t)
        (self._agree_fast if self._fast else self._agree)(err)
        self._tick("agree_wait", t)
        return [res] + later
