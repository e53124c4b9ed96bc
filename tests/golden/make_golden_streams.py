"""Generates the golden STREAM fixtures from the REFERENCE ITSELF. Run in the build container (where
/root/reference exists, after `make -C oracle ref gpu` and `make -C tests/emu`):
    python tests/golden/make_golden_streams.py

  streams/r0_<kind>_<n>_w<bits>.nlzm   output of the pristine reference encoder (oracle/_ref/nlzm_r0
                                       -window:<bits> c) on nlzm_b200.synth.make(kind, n): what
                                       nlzm_codec_decompress must restore
  stream_digests.json                  sha256 + size of the stream the reference's own parser and coder
                                       write when fed by the engine through the host shim
                                       (oracle/_ref/libnlzm_ref_emu.so; emulated engine == CUDA engine,
                                       both bit-exact against the oracle): what nlzm_codec_compress must
                                       emit byte for byte
"""
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nlzm_b200 import synth          # noqa: E402
from oracle import refbind as rb     # noqa: E402

R0 = [("text", 60_000, 15), ("longrange", 90_000, 15), ("mixed", 40_000, 20), ("zeros", 70_000, 15), ("text", 1_500, 15)]
FED = [("text", 150_000, 24), ("longrange", 200_000, 15), ("mixed", 120_000, 20), ("text_drift", 90_000, 16),
       ("zeros", 40_000, 15), ("random", 30_000, 15), ("text", 5, 15), ("text", 1, 15), ("text", 0, 15),
       ("longrange", 300_000, 24)]


def make_input(kind, n):
    import numpy as np
    return synth.make(kind, n) if n else np.zeros(0, np.uint8)


def main():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        src, dst = os.path.join(td, "in.bin"), os.path.join(td, "out.nlzm")
        for kind, n, hb in R0:
            make_input(kind, n).tofile(src)
            path = os.path.join(HERE, "streams", f"r0_{kind}_{n}_w{hb}.nlzm")
            if os.path.exists(path):
                os.remove(path)
            rb.r0_cli(f"-window:{hb}", "c", src, path)
            print("r0", kind, n, hb, os.path.getsize(path), "bytes")
        for kind, n, hb in FED:
            x = make_input(kind, n)
            x.tofile(src)
            if os.path.exists(dst):
                os.remove(dst)
            rb.engine_fed_encode(src, dst, hb, emu=True, block_len=70_000)
            blob = open(dst, "rb").read()
            out[f"{kind}:{n}:{hb}"] = {"sha256": hashlib.sha256(blob).hexdigest(), "size": len(blob),
                                       "input_sha256": hashlib.sha256(x.tobytes()).hexdigest()}
            print("fed", kind, n, hb, len(blob), "bytes")
    json.dump(out, open(os.path.join(HERE, "stream_digests.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
