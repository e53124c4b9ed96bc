"""Times the host pipeline (parser + stream writer) alone on this machine's CPU: candidates come
from the CPU oracle (recorded to a file first), so no GPU is needed. Measurement tool, not product.
    python tools/host_bench.py [kind] [n] [window_bits] [repeats]
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nlzm_b200 import codec, synth          # noqa: E402
from oracle import oracle as orc            # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "text"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
    hb = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    td = os.environ.get("HOST_BENCH_DIR", "/tmp/hb")
    os.makedirs(td, exist_ok=True)
    tag = f"{kind}_{n}_{hb}"
    paths = [os.path.join(td, f"{tag}.{e}") for e in ("txt", "off", "dist", "len")]
    x = synth.make(kind, n)
    if not all(os.path.exists(p) for p in paths):
        orc.build()
        t = time.time()
        off, dist, ln = orc.find(x, hb, orc.F_ALL)
        print(f"oracle candidates: {dist.size} steps in {time.time() - t:.1f} s", file=sys.stderr)
        x.tofile(paths[0])
        off.astype(np.uint64).tofile(paths[1])
        dist.astype(np.uint32).tofile(paths[2])
        ln.astype(np.uint16).tofile(paths[3])
    exe = os.path.join(td, "host_bench")
    subprocess.check_call(["g++", "-O3", "-g", "-std=c++17", "-Wall", os.path.join(ROOT, "tools", "host_bench.cpp"), "-o", exe,
                           "-L" + os.path.join(ROOT, "tests", "emu"), "-lnlzm_mf_emu",
                           "-Wl,-rpath," + os.path.join(ROOT, "tests", "emu")])
    out = os.path.join(td, f"{tag}.nlzm")
    line = subprocess.check_output([exe, *paths, str(hb), out, str(reps)], text=True)
    res = json.loads(line)
    L = codec.bind_prototypes(C.CDLL(os.path.join(ROOT, "tests", "emu", "libnlzm_codec_emu.so")))
    res["roundtrip"] = codec.decompress(open(out, "rb").read(), lib=L) == x.tobytes()
    res["workload"] = f"{kind} {n} -window:{hb}"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
