"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the C restatement of the
reference matchers, oracle/nlzm_oracle.c) plus numpy helpers to compare candidate lists.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")

F_HT2, F_HT3, F_BT4, F_RK256, F_ALL = 1, 2, 4, 8, 15
MATCH_MAX = 264


class Geom(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("hist_bits", "window", "frame_bits", "chunk_size", "feed_size",
                                          "ht2_bits", "ht3_bits", "bt4_bits", "rk_bits")]


class _Steps(C.Structure):
    _fields_ = [("n_pos", C.c_uint64), ("n_steps", C.c_uint64), ("offsets", C.POINTER(C.c_uint64)),
                ("dist", C.POINTER(C.c_uint32)), ("len", C.POINTER(C.c_uint16))]


_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.nlzm_oracle_geometry.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(Geom)]
        L.nlzm_oracle_find.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_Steps)]
        L.nlzm_oracle_find.restype = C.c_int
        L.nlzm_oracle_free.argtypes = [C.POINTER(_Steps)]
        _lib = L
    return _lib


def geometry(flen: int, hist_bits: int) -> Geom:
    g = Geom()
    lib().nlzm_oracle_geometry(flen, hist_bits, C.byref(g))
    return g


def find(x: np.ndarray, hist_bits: int, finder_mask: int = F_ALL, bt_max_tests: int = 0):
    """Returns (offsets[u64, n+1], dist[u32], len[u16]) — the merged staircase of the selected finders."""
    x = np.ascontiguousarray(x, dtype=np.uint8)
    s = _Steps()
    rc = lib().nlzm_oracle_find(x.ctypes.data, x.size, hist_bits, finder_mask, bt_max_tests, C.byref(s))
    if rc:
        raise RuntimeError(f"nlzm_oracle_find rc={rc}")
    try:
        n, m = int(s.n_pos), int(s.n_steps)
        off = np.ctypeslib.as_array(s.offsets, (n + 1,)).copy()
        dist = np.ctypeslib.as_array(s.dist, (m,)).copy() if m else np.zeros(0, np.uint32)
        ln = np.ctypeslib.as_array(s.len, (m,)).copy() if m else np.zeros(0, np.uint16)
    finally:
        lib().nlzm_oracle_free(C.byref(s))
    return off, dist, ln


# ---- helpers shared by the tests -------------------------------------------------------------

def records_to_csr(n_pos: int, pos: np.ndarray, dist: np.ndarray, ln: np.ndarray):
    """Merge an arbitrary multiset of valid candidates (pos, dist, len) into per-position staircases,
    exactly as repeated MatchTable::Update calls would (NLZM.cpp:835-852): a candidate survives iff
    no other candidate at the same position has dist <= its dist and len >= its len."""
    pos = np.asarray(pos, dtype=np.int64)
    dist = np.asarray(dist, dtype=np.int64)
    ln = np.asarray(ln, dtype=np.int64)
    # sort by (pos, dist asc, len desc); keep records whose len exceeds every earlier len in the group
    order = np.lexsort((-ln, dist, pos))
    pos, dist, ln = pos[order], dist[order], ln[order]
    keep = np.ones(pos.size, dtype=bool)
    if pos.size:
        # running max of len within each position group (group-wise cummax via offset trick)
        grp_start = np.r_[True, pos[1:] != pos[:-1]]
        gid = np.cumsum(grp_start) - 1
        big = gid * 1024 + ln            # len < 1024, so groups never interleave
        cm = np.maximum.accumulate(big)
        prev = np.r_[-1, cm[:-1]]
        keep = big > prev
    pos, dist, ln = pos[keep], dist[keep], ln[keep]
    off = np.zeros(n_pos + 1, dtype=np.uint64)
    np.cumsum(np.bincount(pos, minlength=n_pos), out=off[1:])
    return off, dist.astype(np.uint32), ln.astype(np.uint16)


def csr_equal(a, b) -> bool:
    return all(np.array_equal(u, v) for u, v in zip(a, b))


def first_diff(a, b):
    """Human-readable description of the first differing position between two CSR triples."""
    oa, da, la = a
    ob, db, lb = b
    n = min(oa.size, ob.size) - 1
    ca, cb = np.diff(oa[:n + 1].astype(np.int64)), np.diff(ob[:n + 1].astype(np.int64))
    bad = np.flatnonzero(ca != cb)
    cand = [int(bad[0])] if bad.size else []
    m = min(da.size, db.size)
    neq = np.flatnonzero((da[:m] != db[:m]) | (la[:m] != lb[:m]))
    if neq.size:
        cand.append(int(np.searchsorted(oa, neq[0], side="right") - 1))
    if not cand:
        return None
    p = min(cand)
    return (p, list(zip(la[oa[p]:oa[p + 1]].tolist(), da[oa[p]:oa[p + 1]].tolist())),
            list(zip(lb[ob[p]:ob[p + 1]].tolist(), db[ob[p]:ob[p + 1]].tolist())))
