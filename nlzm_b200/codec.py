"""ctypes binding of include/nlzm_codec.h: the host pipeline (parser + stream writer / reader)
around the B200 match-finding engine.

`compress` replaces the reference's `nlzm -window:N c` (encode_file, NLZM.cpp:1711-1910) and emits
the reference's stream format; `decompress` replaces `nlzm d` (decode_file, NLZM.cpp:1912-2039).
compress drives libnlzm_mf and therefore needs a CUDA device: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libnlzm_codec.so")

EXPORTS = ["nlzm_codec_abi_version", "nlzm_codec_compress", "nlzm_codec_decompress", "nlzm_codec_free",
           "nlzm_codec_last_error"]


class CodecConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("window_bits", C.c_uint32), ("device", C.c_int32),
                ("reserved", C.c_uint32), ("block_len", C.c_uint64),
                ("n_devices", C.c_uint32), ("devices", C.c_int32 * 8), ("reserved2", C.c_uint32)]


class CodecStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("in_bytes", "out_bytes", "literals", "matches", "reps", "frames", "parses",
                                          "steps_served", "engine_blocks")] + \
               [("ms_total", C.c_double), ("ms_engine_wait", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def bind_prototypes(L):
    L.nlzm_codec_abi_version.restype = C.c_int
    L.nlzm_codec_compress.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(CodecConfig), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_uint64), C.POINTER(CodecStats)]
    L.nlzm_codec_decompress.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.nlzm_codec_free.argtypes = [C.c_void_p]
    L.nlzm_codec_free.restype = None
    L.nlzm_codec_last_error.restype = C.c_char_p
    return L


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = bind_prototypes(C.CDLL(LIB_PATH))
        if _lib.nlzm_codec_abi_version() != 2:
            raise RuntimeError("libnlzm_codec.so ABI version mismatch")
    return _lib


class CodecError(RuntimeError):
    pass


def _take(L, ptr, n) -> bytes:
    try:
        return C.string_at(ptr.value, n.value) if n.value else b""
    finally:
        L.nlzm_codec_free(ptr)


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def compress(data, window_bits: int = 22, device: int = 0, block_len: int = 0, lib=None, with_stats: bool = False,
             devices=None):
    """bytes / uint8 array -> NLZM stream (bytes). `lib` overrides the library (tests: emulated engine).
    devices: list of CUDA ordinals to spread the engine blocks over (same stream as with one device)."""
    L = lib or load()
    x = _as_u8(data)
    cfg = CodecConfig(C.sizeof(CodecConfig), window_bits, device, 0, block_len)
    if devices:
        cfg.n_devices = len(devices)
        for i, d in enumerate(devices):
            cfg.devices[i] = int(d)
    out, n, st = C.c_void_p(), C.c_uint64(), CodecStats()
    rc = L.nlzm_codec_compress(x.ctypes.data if x.size else None, x.size, C.byref(cfg), C.byref(out), C.byref(n), C.byref(st))
    if rc:
        raise CodecError(f"nlzm_codec_compress rc={rc}: {L.nlzm_codec_last_error().decode()}")
    blob = _take(L, out, n)
    return (blob, st.as_dict()) if with_stats else blob


def decompress(stream, lib=None) -> bytes:
    L = lib or load()
    x = _as_u8(stream)
    out, n = C.c_void_p(), C.c_uint64()
    rc = L.nlzm_codec_decompress(x.ctypes.data if x.size else None, x.size, C.byref(out), C.byref(n))
    if rc:
        raise CodecError(f"nlzm_codec_decompress rc={rc}: {L.nlzm_codec_last_error().decode()}")
    return _take(L, out, n)
