"""End-to-end compress MB/s + ratio (BASELINE.json metric, second half): the reference's own parser,
model and rANS coder fed by the B200 engine through the host shim (oracle/_ref/libnlzm_ref_gpu.so),
next to the pristine reference (oracle/_ref/nlzm_r0) on the same input and -window."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlzm_b200 import synth
from oracle import refbind as rb

kind, n, hb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ref_too = len(sys.argv) > 4 and sys.argv[4] == "ref"
x = synth.make(kind, n)
with tempfile.TemporaryDirectory() as td:
    src, ours, r0, back = (os.path.join(td, f) for f in ("in", "ours.nlzm", "r0.nlzm", "back"))
    x.tofile(src)
    secs, served = rb.engine_fed_encode(src, ours, hb)
    t = time.time(); rb.r0_cli("d", ours, back); td_ = time.time() - t
    ok = open(back, "rb").read() == x.tobytes()
    so = os.path.getsize(ours)
    print(f"{kind} {n} B -window:{hb}: engine-fed encoder {secs:.1f} s = {n/secs/1e6:.2f} MB/s, {so} B (ratio {so/n:.4f}), "
          f"reference decoder round trip {'OK' if ok else 'FAILED'} ({td_:.1f} s), {served} steps consumed")
    if ref_too:
        t = time.time(); rb.r0_cli(f"-window:{hb}", "c", src, r0); tr = time.time() - t
        sr = os.path.getsize(r0)
        print(f"  pristine reference: {tr:.1f} s = {n/tr/1e6:.2f} MB/s, {sr} B; size delta {(so-sr)/sr*100:+.3f} %, speed-up {tr/secs:.1f}x")
