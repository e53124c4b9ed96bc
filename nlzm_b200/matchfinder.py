"""Host-side mirror of the reference's match-finder interface over the C ABI (include/nlzm_mf.h).

The reference encoder owns four finder objects and drives them as
    X.Init(...) / X.FindAndUpdate(mt, hash, P, dict) per position / X.Shift(W) / X.Release()
(NLZM.cpp:1745-1753, 1514-1541, 1786-1792, 1901-1904). `MatchFinders` keeps those names:
Init -> engine creation + input upload, FindAndUpdate -> candidates of a position range,
Shift -> no-op (ring shifts are part of the closed-form geometry), Release -> destroy.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from ._lib import HT2, HT3, BT4, RK256, ALL  # noqa: F401

RAW_STEP_DTYPE = np.dtype([("dist_lo", "<u2"), ("dist_hi", "<u2"), ("len", "<u2")])     # nlzm_mf_step, 6 bytes
STEP_DTYPE = np.dtype([("dist", "<u4"), ("len", "<u2")])                                 # what the wrapper hands out


class MatchFinderError(RuntimeError):
    pass


def geometry(file_len: int, hist_bits: int, lib=None) -> _lib.Geometry:
    L = lib or _lib.load()
    g = _lib.Geometry()
    rc = L.nlzm_mf_get_geometry(file_len, hist_bits, C.byref(g))
    if rc:
        raise MatchFinderError(f"nlzm_mf_get_geometry rc={rc}")
    return g


class MatchFinders:
    """All four finders of one input on one GPU."""

    def __init__(self, lib=None):
        self._L = lib or _lib.load()
        self._h = C.c_void_p()
        self.file_len = 0

    # -- reference-shaped interface -------------------------------------------------------------
    def Init(self, hist_bits: int, data, device: int = 0, finder_mask: int = ALL, max_range: int = 0) -> int:
        """data: numpy uint8 array / bytes (host, copied H2D) or a (device_ptr, nbytes) tuple already in HBM.
        Returns the window size in bytes (the reference's Init returns bytes allocated, used for a printf)."""
        self.Release()
        dev_ptr = None
        if isinstance(data, tuple):
            dev_ptr, n = int(data[0]), int(data[1])
        else:
            data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data,
                                        dtype=np.uint8)
            n = int(data.size)
        cfg = _lib.Config(C.sizeof(_lib.Config), hist_bits, n, device, finder_mask, max_range)
        rc = self._L.nlzm_mf_create(C.byref(cfg), C.byref(self._h))
        if rc:
            self._h = C.c_void_p()
            raise MatchFinderError(f"nlzm_mf_create rc={rc}: {self._L.nlzm_mf_last_error(None).decode()}")
        self.file_len = n
        if dev_ptr is not None:
            rc = self._L.nlzm_mf_set_input_device(self._h, C.c_void_p(dev_ptr), n)
        else:
            rc = self._L.nlzm_mf_set_input(self._h, data.ctypes.data_as(C.c_void_p), n)
        self._check(rc, "set_input")
        return geometry(n, hist_bits, self._L).window

    def FindAndUpdate(self, begin: int = 0, end: int | None = None, slot: int = 0, copy: bool = True):
        """Candidates of positions [begin, end) as (offsets u32[end-begin+1], steps STEP_DTYPE[n]).
        Each step is one MatchTable::Update(dist, len) the reference finders would have issued.
        copy=False returns zero-copy views of the engine's pinned buffers instead (raw 6-byte records)."""
        end = self.file_len if end is None else end
        v = _lib.View()
        self._check(self._L.nlzm_mf_find(self._h, begin, end, slot, C.byref(v)), "find")
        return self._view_to_numpy(v, copy)

    def Shift(self, shift: int) -> None:  # NLZM.cpp:1786-1792: nothing to do, see module docstring
        return None

    def Release(self) -> None:
        if self._h:
            self._L.nlzm_mf_destroy(self._h)
            self._h = C.c_void_p()

    # -- extras -----------------------------------------------------------------------------------
    def find_device(self, begin: int = 0, end: int | None = None, slot: int = 0) -> _lib.View:
        """Same as FindAndUpdate but leaves the result in HBM (view holds device pointers)."""
        end = self.file_len if end is None else end
        v = _lib.View()
        self._check(self._L.nlzm_mf_find_device(self._h, begin, end, slot, C.byref(v)), "find_device")
        return v

    def submit(self, begin: int, end: int, slot: int) -> None:
        self._check(self._L.nlzm_mf_submit(self._h, begin, end, slot), "submit")

    def fetch(self, slot: int, copy: bool = True):
        v = _lib.View()
        self._check(self._L.nlzm_mf_fetch(self._h, slot, C.byref(v)), "fetch")
        return self._view_to_numpy(v, copy)

    # -- position sharding across engines / GPUs (include/nlzm_mf.h: segments) ---------------------------
    def prepare(self, begin: int, end: int) -> None:
        """rank + merge [begin, end) alone; other engines can then import its segments, find() continues from here"""
        self._check(self._L.nlzm_mf_prepare(self._h, begin, end), "prepare")

    def export_segments(self) -> list:
        n = C.c_uint32(0)
        self._check(self._L.nlzm_mf_export_segments(self._h, None, 0, C.byref(n)), "export_segments")
        arr = (_lib.SegmentDesc * max(n.value, 1))()
        self._check(self._L.nlzm_mf_export_segments(self._h, arr, n.value, C.byref(n)), "export_segments")
        return [arr[i] for i in range(n.value)]

    def publish_segments(self, from_pos: int) -> list:
        """copies of the own segments at positions >= from_pos in the engine's export buffer (one IPC handle)"""
        n = C.c_uint32(0)
        self._check(self._L.nlzm_mf_publish_segments(self._h, from_pos, None, 0, C.byref(n)), "publish_segments")
        arr = (_lib.SegmentDesc * max(n.value, 1))()
        self._check(self._L.nlzm_mf_publish_segments(self._h, from_pos, arr, n.value, C.byref(n)), "publish_segments")
        return [arr[i] for i in range(n.value)]

    def import_segment(self, desc, via: int = 0, host_copy=None) -> None:
        """desc: a SegmentDesc from export_segments() of another engine, or its bytes (from another process).
        via 0: same process (peer copy), 1: CUDA IPC handles, 2: host_copy = (elems, ptrs) numpy byte arrays"""
        if isinstance(desc, (bytes, bytearray)):
            desc = _lib.SegmentDesc.from_buffer_copy(desc)
        if host_copy is not None:
            via = 2
            d = _lib.SegmentDesc.from_buffer_copy(bytes(desc))
            d.elems_alloc = host_copy[0].ctypes.data
            d.ptrs_alloc = host_copy[1].ctypes.data
            desc = d
        self._check(self._L.nlzm_mf_import_segment(self._h, C.byref(desc), int(via)), "import_segment")

    def read_segment(self, index: int, desc):
        """host copies (elems, ptrs) of retained segment `index` (desc = export_segments()[index])"""
        el = np.empty(int(desc.elems_bytes), np.uint8)
        pt = np.empty(int(desc.ptrs_bytes), np.uint8)
        self._check(self._L.nlzm_mf_read_segment(self._h, index, el.ctypes.data, pt.ctypes.data), "read_segment")
        return el, pt

    def trim_segments(self, from_pos: int) -> None:
        """keep only positions >= from_pos in the retained segments (what the next shard can reach)"""
        self._check(self._L.nlzm_mf_trim_segments(self._h, from_pos), "trim_segments")

    def drop_segments(self) -> None:
        self._check(self._L.nlzm_mf_drop_segments(self._h), "drop_segments")

    def set_option(self, key: str, value: int) -> None:
        """tuning / test knobs, see nlzm_mf_set_option in include/nlzm_mf.h"""
        self._check(self._L.nlzm_mf_set_option(self._h, key.encode(), value), "set_option")

    def stats(self) -> _lib.Stats:
        s = _lib.Stats()
        self._L.nlzm_mf_get_stats(self._h, C.byref(s))
        return s

    def _view_to_numpy(self, v, copy):
        n = int(v.end - v.begin)
        off = np.ctypeslib.as_array(C.cast(v.offsets, C.POINTER(C.c_uint32)), (n + 1,))
        m = int(v.n_steps)
        if m:
            buf = (C.c_char * (m * RAW_STEP_DTYPE.itemsize)).from_address(v.steps)
            raw = np.frombuffer(buf, dtype=RAW_STEP_DTYPE)
        else:
            raw = np.zeros(0, dtype=RAW_STEP_DTYPE)
        if not copy:
            return off, raw                 # zero-copy view of the pinned buffers (fields dist_lo, dist_hi, len)
        return off.copy(), unpack_steps(raw)

    def _check(self, rc, what):
        if rc:
            msg = self._L.nlzm_mf_last_error(self._h).decode() if self._h else ""
            raise MatchFinderError(f"nlzm_mf_{what} rc={rc}: {msg}")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Release()

    def __del__(self):
        try:
            self.Release()
        except Exception:
            pass


def unpack_steps(raw: np.ndarray) -> np.ndarray:
    """raw 6-byte records (copy=False views) -> STEP_DTYPE array with a 32-bit `dist` field"""
    steps = np.empty(raw.size, dtype=STEP_DTYPE)
    steps["dist"] = raw["dist_lo"].astype(np.uint32) | ((raw["dist_hi"].astype(np.uint32) & 0x0FFF) << 16)
    steps["len"] = raw["len"] & 0x1FF          # bits 9..14 carry the distance slot, dist_hi bits 12..13 the shortest length
    return steps


def unpack_prepricing(raw: np.ndarray):
    """(distance slot, shortest length) the engine computed for every step"""
    return ((raw["len"] >> 9) & 0x3F).astype(np.uint8), (2 + ((raw["dist_hi"] >> 12) & 3)).astype(np.uint8)


def profile(enable: bool, lib=None) -> None:
    (lib or _lib.load()).nlzm_mf_profile(int(enable))


def kernel_times(lib=None) -> dict:
    """{kernel name: (launches, total ms)} accumulated since profile(True)."""
    L = lib or _lib.load()
    n = C.c_uint32(0)
    L.nlzm_mf_get_kernel_times(None, 0, C.byref(n))
    arr = (_lib.KernelTime * max(n.value, 1))()
    L.nlzm_mf_get_kernel_times(arr, n.value, C.byref(n))
    return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].ms)) for i in range(n.value)}
