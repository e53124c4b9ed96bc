"""Generates the golden fixtures from the REFERENCE ITSELF (oracle/_ref/libnlzm_ref.so, built by
oracle/Makefile from /root/reference/NLZM.cpp with the BT4 cap lifted and the skip rule off = "R2").
Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py

Two kinds of fixture:
  small_*.npz   input bytes + the per-finder step records of the hooked REAL encoder (encode_file)
  digests.json  sha256 of the merged candidate CSR of larger seeded synthetic inputs (input is
                regenerated from nlzm_b200.synth, so only the digest is stored)
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nlzm_b200 import synth          # noqa: E402
from oracle import oracle as orc     # noqa: E402
from oracle import refbind as rb     # noqa: E402

SMALL = [("text", 48_000, 15, 1), ("longrange", 80_000, 15, 2), ("mixed", 40_000, 15, 3),
         ("text", 150_000, 17, 4), ("longrange", 100_000, 24, 5)]
DIGEST = [("text", 1_000_000, 24), ("text", 1_500_000, 15), ("text_drift", 1_200_000, 20),
          ("longrange", 2_000_000, 24), ("longrange", 1_000_000, 16), ("mixed", 800_000, 24),
          ("mixed", 700_000, 17), ("random", 300_000, 22), ("zeros", 200_000, 18)]


def digest(csr):
    h = hashlib.sha256()
    for a in csr:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    with tempfile.TemporaryDirectory() as td:
        for kind, n, hb, seed in SMALL:
            x = synth.make(kind, n, seed)
            src, dst = os.path.join(td, "in.bin"), os.path.join(td, "out.nlzm")
            x.tofile(src)
            recs = rb.encode_dump(src, dst, hb, mode=rb.R2)          # the real encoder, hooked
            size = os.path.getsize(dst)
            np.savez_compressed(os.path.join(HERE, f"small_{kind}_{n}_w{hb}.npz"), x=x, hist_bits=hb,
                                pos=recs["pos"].astype(np.uint32), dist=recs["dist"], len=recs["len"],
                                finder=recs["finder"].astype(np.uint8), r2_size=size)
            print("small", kind, n, hb, recs.size, "records, R2 stream", size, "bytes")
    out = {}
    for kind, n, hb in DIGEST:
        x = synth.make(kind, n)
        recs, _ = rb.matchfind(x, hb, mode=rb.R2)                    # reference finder objects, R2
        csr = orc.records_to_csr(x.size, recs["pos"], recs["dist"], recs["len"])
        csr = (csr[0].astype(np.uint64), csr[1].astype(np.uint32), csr[2].astype(np.uint16))
        out[f"{kind}:{n}:{hb}"] = {"sha256": digest(csr), "steps": int(csr[1].size),
                                   "input_sha256": hashlib.sha256(x.tobytes()).hexdigest()}
        print("digest", kind, n, hb, csr[1].size)
    json.dump(out, open(os.path.join(HERE, "digests.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
