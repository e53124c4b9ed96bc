"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libnlzm_ref.so (the reference itself,
compiled by oracle/Makefile from /root/reference/NLZM.cpp with dump hooks).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libnlzm_ref.so")
REF_R0 = os.path.join(_HERE, "_ref", "nlzm_r0")

REC_DTYPE = np.dtype([("pos", "<u8"), ("dist", "<u4"), ("len", "<u2"), ("finder", "<u2")])
FINDERS = {"ht2": 0, "ht3": 1, "bt4": 2, "rk256": 3, "final": 4}

R0 = (256, 64)            # as shipped (NLZM.cpp:777, 734)
R1 = (0x7FFFFFFF, 64)     # BT4 cap lifted
R2 = (0x7FFFFFFF, 1000)   # cap lifted + skip rule off: finder output is a pure function of the bytes


def available() -> bool:
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(REF_SO)
        L.ref_set_mode.argtypes = [C.c_uint32, C.c_uint32]
        L.ref_set_dump.argtypes = [C.c_int]
        L.ref_dump_count.restype = C.c_uint64
        L.ref_dump_data.restype = C.c_void_p
        L.ref_encode.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32]
        L.ref_decode.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_matchfind.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.ref_matchfind.restype = C.c_double
        _lib = L
    return _lib


def _take_dump() -> np.ndarray:
    L = lib()
    n = L.ref_dump_count()
    if n == 0:
        return np.zeros(0, dtype=REC_DTYPE)
    buf = (C.c_char * (n * REC_DTYPE.itemsize)).from_address(L.ref_dump_data())
    out = np.frombuffer(buf, dtype=REC_DTYPE).copy()
    L.ref_dump_clear()
    return out


def encode_dump(in_path: str, out_path: str, hist_bits: int, mode=R2, dump_mask: int = 0xF) -> np.ndarray:
    """Run the reference's real encoder (hooked) and return every finder step it produced."""
    L = lib()
    L.ref_set_mode(*mode)
    L.ref_dump_clear()
    L.ref_set_dump(dump_mask)
    rc = L.ref_encode(in_path.encode(), out_path.encode(), hist_bits)
    L.ref_set_dump(0)
    if rc:
        raise RuntimeError(f"ref_encode rc={rc}")
    return _take_dump()


def decode(in_path: str, out_path: str) -> None:
    rc = lib().ref_decode(in_path.encode(), out_path.encode())
    if rc:
        raise RuntimeError(f"ref_decode rc={rc}")


def matchfind(x: np.ndarray, hist_bits: int, finder_mask: int = 0xF, mode=R2, use_carry: bool = False,
              dump_mask: int = 0xF):
    """Matcher-only driver around the reference's finder objects. Returns (records, seconds)."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.uint8)
    # the reference reads VALUE4 at the last positions: keep a readable tail
    buf = np.zeros(x.size + 16, dtype=np.uint8)
    buf[:x.size] = x
    L.ref_set_mode(*mode)
    L.ref_dump_clear()
    L.ref_set_dump(dump_mask)
    upd = C.c_uint64(0)
    secs = L.ref_matchfind(buf.ctypes.data, x.size, hist_bits, finder_mask, int(use_carry), C.byref(upd))
    L.ref_set_dump(0)
    return _take_dump(), secs


# ---- the reference's parser + coder fed by the engine through the host shim (integration demo) ----
REF_GPU_SO = os.path.join(_HERE, "_ref", "libnlzm_ref_gpu.so")
REF_EMU_SO = os.path.join(_HERE, "_ref", "libnlzm_ref_emu.so")


def engine_fed_encode(in_path: str, out_path: str, hist_bits: int, emu: bool = False, device: int = 0,
                      block_len: int = 0):
    """Returns (seconds, steps_served). emu=True links the sequential emulation (CPU-only check)."""
    path = REF_EMU_SO if emu else REF_GPU_SO
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run `make -C oracle gpu` where /root/reference exists")
    L = C.CDLL(path)
    L.refgpu_encode.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_int, C.c_uint64, C.POINTER(C.c_double),
                                C.POINTER(C.c_uint64)]
    secs, served = C.c_double(0), C.c_uint64(0)
    rc = L.refgpu_encode(in_path.encode(), out_path.encode(), hist_bits, device, block_len, C.byref(secs), C.byref(served))
    if rc:
        raise RuntimeError(f"refgpu_encode rc={rc}")
    return secs.value, served.value


def r0_cli(*args) -> str:
    """Run the pristine reference binary (oracle/_ref/nlzm_r0)."""
    import subprocess
    return subprocess.run([REF_R0, *args], check=True, capture_output=True, text=True).stdout
