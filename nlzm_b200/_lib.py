"""ctypes prototypes of include/nlzm_mf.h and the loader of the CUDA library.

The product loads nlzm_b200/csrc/libnlzm_mf.so (built by nlzm_b200.build / __graft_entry__.build)
and nothing else: there is no CPU fallback. A missing library or a machine without a CUDA device
raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NLZM_MF_LIB") or os.path.join(HERE, "csrc", "libnlzm_mf.so")   # override: A/B builds (tools/)

HT2, HT3, BT4, RK256, ALL = 1, 2, 4, 8, 15


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("hist_bits", C.c_uint32), ("file_len", C.c_uint64),
                ("device", C.c_int32), ("finder_mask", C.c_uint32), ("max_range", C.c_uint64)]


class Geometry(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("hist_bits", "window", "frame_bits", "chunk_size", "feed_size",
                                          "ht2_bits", "ht3_bits", "bt4_bits", "rk_bits")]


class View(C.Structure):
    _fields_ = [("begin", C.c_uint64), ("end", C.c_uint64), ("n_steps", C.c_uint64),
                ("offsets", C.c_void_p), ("steps", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("tuples_last", C.c_uint64)] + \
               [(n, C.c_float) for n in ("ms_rank", "ms_levels", "ms_ht", "ms_rk", "ms_merge", "ms_total", "ms_d2h",
                                          "ms_cross", "ms_prepare")] + \
               [("segments_queried", C.c_uint32), ("segments_retained", C.c_uint32), ("ms_import", C.c_float),
                ("reserved", C.c_uint32), ("bytes_imported", C.c_uint64)]


class SegmentDesc(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("pos_begin", "pos_end", "origin", "n_elems", "elems_offset_bytes", "elems_bytes",
                                          "ptrs_offset_bytes", "ptrs_bytes")] + \
               [("elems_alloc", C.c_void_p), ("ptrs_alloc", C.c_void_p), ("device", C.c_int32), ("flags", C.c_uint32),
                ("ipc_elems", C.c_uint8 * 64), ("ipc_ptrs", C.c_uint8 * 64)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 56), ("launches", C.c_uint64), ("ms", C.c_double)]


EXPORTS = ["nlzm_mf_set_option", "nlzm_mf_profile", "nlzm_mf_get_kernel_times", "nlzm_mf_abi_version", "nlzm_mf_get_geometry", "nlzm_mf_create", "nlzm_mf_destroy",
           "nlzm_mf_last_error", "nlzm_mf_set_input", "nlzm_mf_set_input_device", "nlzm_mf_find",
           "nlzm_mf_find_device", "nlzm_mf_submit", "nlzm_mf_fetch", "nlzm_mf_get_stats",
           "nlzm_mf_prepare", "nlzm_mf_export_segments", "nlzm_mf_import_segment", "nlzm_mf_drop_segments", "nlzm_mf_read_segment", "nlzm_mf_trim_segments", "nlzm_mf_publish_segments"]


def bind_prototypes(L):
    """Attach argtypes/restypes for every symbol include/nlzm_mf.h declares."""
    L.nlzm_mf_abi_version.restype = C.c_int
    L.nlzm_mf_get_geometry.argtypes = [C.c_uint64, C.c_uint32, C.POINTER(Geometry)]
    L.nlzm_mf_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.nlzm_mf_destroy.argtypes = [C.c_void_p]
    L.nlzm_mf_destroy.restype = None
    L.nlzm_mf_last_error.argtypes = [C.c_void_p]
    L.nlzm_mf_last_error.restype = C.c_char_p
    L.nlzm_mf_set_input.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.nlzm_mf_set_input_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    for name in ("nlzm_mf_find", "nlzm_mf_find_device"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(View)]
    L.nlzm_mf_submit.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
    L.nlzm_mf_fetch.argtypes = [C.c_void_p, C.c_int, C.POINTER(View)]
    L.nlzm_mf_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.nlzm_mf_prepare.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.nlzm_mf_export_segments.argtypes = [C.c_void_p, C.POINTER(SegmentDesc), C.c_uint32, C.POINTER(C.c_uint32)]
    L.nlzm_mf_import_segment.argtypes = [C.c_void_p, C.POINTER(SegmentDesc), C.c_int]
    L.nlzm_mf_drop_segments.argtypes = [C.c_void_p]
    L.nlzm_mf_publish_segments.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(SegmentDesc), C.c_uint32, C.POINTER(C.c_uint32)]
    L.nlzm_mf_trim_segments.argtypes = [C.c_void_p, C.c_uint64]
    L.nlzm_mf_read_segment.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.nlzm_mf_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
    L.nlzm_mf_profile.argtypes = [C.c_int]
    L.nlzm_mf_get_kernel_times.argtypes = [C.POINTER(KernelTime), C.c_uint32, C.POINTER(C.c_uint32)]
    return L


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = bind_prototypes(C.CDLL(LIB_PATH))
        if _lib.nlzm_mf_abi_version() != 3:
            raise RuntimeError("libnlzm_mf.so ABI version mismatch")
    return _lib
