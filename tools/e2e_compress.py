"""End-to-end compress MB/s + ratio (BASELINE.json metric, second half), three ways on the same input
and -window:
  * this repo's own host pipeline (nlzm_b200.codec.compress: libnlzm_codec over the B200 engine);
  * the reference's own parser, model and rANS coder fed by the engine through the host shim
    (oracle/_ref/libnlzm_ref_gpu.so) — must be byte-identical to the first;
  * with `ref`: the pristine reference (oracle/_ref/nlzm_r0).
    python tools/e2e_compress.py <kind> <n> <window_bits> [ref] [nofed]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlzm_b200 import synth
from oracle import refbind as rb

kind, n, hb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ref_too = "ref" in sys.argv[4:]
fed_too = "nofed" not in sys.argv[4:]
import json
from nlzm_b200 import codec
x = synth.make(kind, n)
with tempfile.TemporaryDirectory() as td:
    src, ours, r0, back = (os.path.join(td, f) for f in ("in", "ours.nlzm", "r0.nlzm", "back"))
    x.tofile(src)
    codec.compress(x[:1 << 20], hb)                                   # warm-up: CUDA context, library load
    t = time.time(); blob, st = codec.compress(x, hb, with_stats=True); tc = time.time() - t
    t = time.time(); okc = codec.decompress(blob) == x.tobytes(); tdc = time.time() - t
    print(f"{kind} {n} B -window:{hb}: own pipeline {tc:.2f} s = {n/tc/1e6:.2f} MB/s, {len(blob)} B (ratio {len(blob)/n:.4f}), "
          f"engine wait {st['ms_engine_wait']:.0f} ms in {st['engine_blocks']} blocks, own decoder {n/tdc/1e6:.1f} MB/s "
          f"round trip {'OK' if okc else 'FAILED'}")
    print(json.dumps({"workload": f"{kind} {n} -window:{hb}", "compress_MBps": round(n/tc/1e6, 3), "stream_bytes": len(blob),
                      "decompress_MBps": round(n/tdc/1e6, 2), "roundtrip": okc, **st}))
    if not fed_too:
        sys.exit(0 if okc else 1)
    secs, served = rb.engine_fed_encode(src, ours, hb)
    print(f"  own stream == engine-fed reference encoder stream: {blob == open(ours, 'rb').read()}")
    t = time.time(); rb.r0_cli("d", ours, back); td_ = time.time() - t
    ok = open(back, "rb").read() == x.tobytes()
    so = os.path.getsize(ours)
    print(f"{kind} {n} B -window:{hb}: engine-fed encoder {secs:.1f} s = {n/secs/1e6:.2f} MB/s, {so} B (ratio {so/n:.4f}), "
          f"reference decoder round trip {'OK' if ok else 'FAILED'} ({td_:.1f} s), {served} steps consumed")
    if ref_too:
        t = time.time(); rb.r0_cli(f"-window:{hb}", "c", src, r0); tr = time.time() - t
        sr = os.path.getsize(r0)
        print(f"  pristine reference: {tr:.1f} s = {n/tr/1e6:.2f} MB/s, {sr} B; size delta {(so-sr)/sr*100:+.3f} %, speed-up {tr/secs:.1f}x")
