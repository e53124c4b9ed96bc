"""BASELINE-scale golden digests from the REFERENCE ITSELF (oracle/_ref/libnlzm_ref.so, R2 mode: BT4 cap
lifted, skip rule off). Run in the build container (where /root/reference exists), takes tens of minutes:

    python tests/golden/make_golden_big.py [key ...]

The reference's finder objects are driven over the whole input (ref_matchfind, all four finders, no
carry); every finder's staircase and the merged table are streamed to scratch files (the record list
of 100 MB of text would not fit in memory) and digested:

    digest = sha256( sha256(counts u8[n]) + sha256(dist u32[m]) + sha256(len u16[m]) )     (hex, concatenated)

counts[a] = number of staircase steps of position a. tests/test_gpu_scale.py computes the same digest from the
engine's CSR output, per finder mask and for all four finders.
"""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nlzm_b200 import synth          # noqa: E402
from oracle import refbind as rb     # noqa: E402

# (kind, n, hist_bits): C2 in full; two inputs that keep -window:28 after the reference's shrink rule
BIG = [("text", 100_000_000, 24), ("longrange", 140_000_000, 28), ("text_drift", 140_000_000, 28)]
OUT = os.path.join(HERE, "digests_big.json")
FINDER_KEYS = ["ht2", "ht3", "bt4", "rk256", "all"]


def file_sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def combine(c, d, l):
    return hashlib.sha256((c + d + l).encode()).hexdigest()


def csr_digest(offsets, dist, ln):
    """same digest from a CSR triple (used by the GPU tests)"""
    cnt = np.diff(offsets.astype(np.int64)).astype(np.uint8)
    return combine(hashlib.sha256(cnt.tobytes()).hexdigest(),
                   hashlib.sha256(np.ascontiguousarray(dist, dtype=np.uint32).tobytes()).hexdigest(),
                   hashlib.sha256(np.ascontiguousarray(ln, dtype=np.uint16).tobytes()).hexdigest())


def main():
    want = set(sys.argv[1:])
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    L = rb.lib()
    L.ref_set_stream.argtypes = [C.c_char_p]
    for kind, n, hb in BIG:
        key = f"{kind}:{n}:{hb}"
        if want and key not in want:
            continue
        x = synth.make(kind, n)
        buf = np.zeros(x.size + 16, dtype=np.uint8)
        buf[:x.size] = x
        with tempfile.TemporaryDirectory(dir=os.environ.get("NLZM_SCRATCH", "/tmp")) as td:
            prefix = os.path.join(td, "dump")
            L.ref_set_mode(*rb.R2)
            L.ref_set_dump(0)
            assert L.ref_set_stream(prefix.encode()) == 0
            t = time.time()
            upd = C.c_uint64(0)
            secs = L.ref_matchfind(buf.ctypes.data, x.size, hb, 15, 0, C.byref(upd))
            L.ref_set_stream(None)
            entry = {"input_sha256": hashlib.sha256(x.tobytes()).hexdigest(), "reference_seconds": round(secs, 1),
                     "mode": "R2 (cap lifted, skip rule off), oracle/_ref/libnlzm_ref.so ref_matchfind"}
            for f, name in enumerate(FINDER_KEYS):
                c, d, l = (f"{prefix}.f{f}.{e}" for e in ("cnt", "dist", "len"))
                assert os.path.getsize(c) == n, (name, os.path.getsize(c))
                entry[name] = {"steps": os.path.getsize(d) // 4, "sha256": combine(file_sha(c), file_sha(d), file_sha(l))}
            out[key] = entry
            print(key, entry, f"{time.time() - t:.0f}s", flush=True)
        json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
