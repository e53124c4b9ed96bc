"""Host pipeline: nlzm_codec_compress / nlzm_codec_decompress.

Parity bar, same as the matcher's: byte-exact.
  * compress == the stream the reference's own parser + coder write when fed by the engine
    (golden sha256 made from oracle/_ref/libnlzm_ref_emu.so, and live against oracle/_ref where it is
    built), and the pristine reference decoder restores the input from it;
  * decompress restores the input from streams written by the pristine reference encoder
    (tests/golden/streams/r0_*.nlzm).
CPU tests run the engine's sequential emulation (tests/emu); `-m gpu` tests run the CUDA engine."""
import ctypes as C
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
DIGESTS = json.load(open(os.path.join(GOLD, "stream_digests.json")))


def _input(kind, n):
    from nlzm_b200 import synth
    return synth.make(kind, n) if n else np.zeros(0, np.uint8)


@pytest.fixture(scope="session")
def codec_emu(emu_lib):
    from nlzm_b200 import codec
    return codec.bind_prototypes(C.CDLL(os.path.join(ROOT, "tests", "emu", "libnlzm_codec_emu.so")))


@pytest.fixture(scope="session")
def codec_cuda(cuda_lib):
    from nlzm_b200 import codec
    return codec.load()


def test_codec_library_exports_every_declared_symbol():
    """No compute: the library loads and has every entry point include/nlzm_codec.h declares."""
    import re
    from nlzm_b200 import build, codec
    build.build_codec()
    L = C.CDLL(codec.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "nlzm_codec.h")).read()
    declared = sorted(set(re.findall(r"\b(nlzm_codec_[a-z_]+)\s*\(", header)))
    assert declared == sorted(codec.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.nlzm_codec_abi_version() == 2
    assert C.sizeof(codec.CodecConfig) == 64 and C.sizeof(codec.CodecStats) == 88


def test_compress_without_a_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from nlzm_b200 import build, codec
    build.build_codec()
    with pytest.raises(codec.CodecError, match="no CUDA device|no CPU fallback"):
        codec.compress(b"there is no host matcher to fall back to")


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "streams", "r0_*.nlzm"))))
def test_decompress_restores_reference_encoder_streams(codec_emu, name):
    from nlzm_b200 import codec
    _, kind, n, w = name[:-5].rsplit("_", 3)
    stream = open(os.path.join(GOLD, "streams", name), "rb").read()
    assert codec.decompress(stream, lib=codec_emu) == _input(kind, int(n)).tobytes()


@pytest.mark.parametrize("key", sorted(k for k in DIGESTS if int(k.split(":")[1]) <= 200_000))
def test_compress_matches_engine_fed_reference_stream_emulated(codec_emu, key):
    from nlzm_b200 import codec
    kind, n, hb = key.split(":")
    x = _input(kind, int(n))
    assert hashlib.sha256(x.tobytes()).hexdigest() == DIGESTS[key]["input_sha256"]
    blob, st = codec.compress(x, int(hb), block_len=50_000, lib=codec_emu, with_stats=True)
    assert len(blob) == DIGESTS[key]["size"]
    assert hashlib.sha256(blob).hexdigest() == DIGESTS[key]["sha256"]
    assert codec.decompress(blob, lib=codec_emu) == x.tobytes()
    assert st["in_bytes"] == x.size and st["out_bytes"] == len(blob)
    assert st["literals"] + st["matches"] + st["reps"] > 0 or x.size == 0


def test_compress_short_first_block_emulated(codec_emu):
    """With a large block length the first engine block is cut short (parsing starts early); the
    stream must not depend on where blocks are cut."""
    from nlzm_b200 import codec
    key = "text:150000:24"
    blob, st = codec.compress(_input("text", 150_000), 24, block_len=600_000, lib=codec_emu, with_stats=True)
    assert st["engine_blocks"] == 2
    assert hashlib.sha256(blob).hexdigest() == DIGESTS[key]["sha256"]


def test_compress_live_against_reference_parser_and_decoder_emulated(tmp_path, codec_emu):
    """Same comparison made live where oracle/_ref is built: identical to the engine-fed reference
    encoder, restored by the pristine reference decoder. Block length must not matter."""
    from oracle import refbind as rb
    from nlzm_b200 import codec
    for path in (rb.REF_EMU_SO, rb.REF_R0):
        if not os.path.exists(path):
            pytest.skip("oracle/_ref not built (needs /root/reference)")
    x = _input("mixed", 70_000)
    src, ref, ours, back = (str(tmp_path / f) for f in ("in.bin", "ref.nlzm", "ours.nlzm", "back.bin"))
    x.tofile(src)
    rb.engine_fed_encode(src, ref, 16, emu=True, block_len=33_000)
    blob = codec.compress(x, 16, block_len=21_000, lib=codec_emu)
    assert blob == open(ref, "rb").read()
    open(ours, "wb").write(blob)
    rb.r0_cli("d", ours, back)
    assert open(back, "rb").read() == x.tobytes()


def test_decompress_rejects_damaged_streams(codec_emu):
    from nlzm_b200 import codec
    good = open(os.path.join(GOLD, "streams", "r0_text_60000_w15.nlzm"), "rb").read()
    for bad in (b"", good[:3], good[:7], good[:200], good[:-4], good[:len(good) // 2], b"\x00\x63" + good[2:]):
        with pytest.raises(codec.CodecError):
            codec.decompress(bad, lib=codec_emu)
    # flipped payload bytes must never crash; they either fail cleanly or decode to different bytes
    rng = np.random.default_rng(5)
    for _ in range(40):
        b = bytearray(good)
        b[int(rng.integers(4, len(b)))] ^= 1 << int(rng.integers(0, 8))
        try:
            codec.decompress(bytes(b), lib=codec_emu)
        except codec.CodecError:
            pass


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_compress_matches_engine_fed_reference_stream_gpu(codec_cuda, key):
    from nlzm_b200 import codec
    kind, n, hb = key.split(":")
    x = _input(kind, int(n))
    blob = codec.compress(x, int(hb), block_len=1 << 16)
    assert len(blob) == DIGESTS[key]["size"]
    assert hashlib.sha256(blob).hexdigest() == DIGESTS[key]["sha256"]
    assert codec.decompress(blob) == x.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,hb", [("text", 6_000_000, 24), ("longrange", 5_000_000, 22), ("text", 2_000_000, 15)])
def test_compress_live_against_reference_parser_and_decoder_gpu(tmp_path, codec_cuda, kind, n, hb):
    from oracle import refbind as rb
    from nlzm_b200 import codec
    x = _input(kind, n)
    blob, st = codec.compress(x, hb, with_stats=True)
    assert codec.decompress(blob) == x.tobytes()
    print(f"{kind} {n} -window:{hb}: {len(blob)} B, {st['ms_total']:.0f} ms ({n / st['ms_total'] / 1e3:.2f} MB/s), "
          f"engine wait {st['ms_engine_wait']:.0f} ms, {st['engine_blocks']} blocks")
    if not (os.path.exists(rb.REF_GPU_SO) and os.path.exists(rb.REF_R0)):
        pytest.skip("oracle/_ref not built; round trip only")
    src, ref, ours, back = (str(tmp_path / f) for f in ("in.bin", "ref.nlzm", "ours.nlzm", "back.bin"))
    x.tofile(src)
    rb.engine_fed_encode(src, ref, hb, emu=False, block_len=1 << 21)
    assert blob == open(ref, "rb").read()
    open(ours, "wb").write(blob)
    rb.r0_cli("d", ours, back)
    assert open(back, "rb").read() == x.tobytes()


@pytest.mark.parametrize("kind", ["ab", "abc_runs", "period", "zeros_ones", "words"])
def test_compress_fuzz_against_reference_parser_emulated(tmp_path, codec_emu, kind):
    """Small adversarial inputs (tiny alphabets, periods, runs, lengths around the 264-byte cap and the
    4096-byte segment cap, several chunks): own parser + coder vs the reference's, same emulated engine."""
    from oracle import refbind as rb
    from nlzm_b200 import codec
    from test_fuzz import _gen
    for path in (rb.REF_EMU_SO, rb.REF_R0):
        if not os.path.exists(path):
            pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(sum(kind.encode()) + 1)
    src, ref = str(tmp_path / "in.bin"), str(tmp_path / "ref.nlzm")
    for n in (2, 3, 4, 7, 263, 264, 265, 266, 700, 4095, 4096, 4097, 4400, 14_848, 14_849, 31_000):
        x = _gen(kind, n, rng)
        x.tofile(src)
        if os.path.exists(ref):
            os.remove(ref)
        rb.engine_fed_encode(src, ref, 15, emu=True, block_len=20_000)
        blob = codec.compress(x, 15, block_len=9_000, lib=codec_emu)
        assert blob == open(ref, "rb").read(), (kind, n)
        assert codec.decompress(blob, lib=codec_emu) == x.tobytes(), (kind, n)


def test_compress_is_reentrant_emulated(codec_emu):
    """Independent inputs may be compressed from several host threads at once (one engine instance
    each): the sequential parse does not scale inside one stream, replicas across streams do."""
    import threading
    from nlzm_b200 import codec
    keys = ["mixed:120000:20", "text_drift:90000:16", "zeros:40000:15", "random:30000:15"]
    got = {}

    def work(key):
        kind, n, hb = key.split(":")
        got[key] = codec.compress(_input(kind, int(n)), int(hb), block_len=40_000, lib=codec_emu)

    threads = [threading.Thread(target=work, args=(k,)) for k in keys]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in keys:
        assert hashlib.sha256(got[k]).hexdigest() == DIGESTS[k]["sha256"], k


@pytest.mark.parametrize("kind,n,hb,ndev", [("text", 260_000, 15, 2), ("longrange", 300_000, 16, 3), ("mixed", 200_000, 15, 4)])
def test_compress_multi_engine_feed_same_stream_emulated(codec_emu, kind, n, hb, ndev):
    """MultiBlockFeed: blocks dealt round-robin to several engines of one process, the window behind a block copied
    from the engine that owns the block before it; the stream is byte-identical to the one-engine path"""
    from nlzm_b200 import codec
    x = _input(kind, n)
    one = codec.compress(x, hb, lib=codec_emu, block_len=40_000)
    many, st = codec.compress(x, hb, lib=codec_emu, block_len=40_000, devices=list(range(ndev)), with_stats=True)
    assert many == one
    assert st["engine_blocks"] >= 3
    assert codec.decompress(many, lib=codec_emu) == x.tobytes()


@pytest.mark.gpu
def test_compress_two_gpus_same_stream():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from nlzm_b200 import codec, synth
    x = np.concatenate([synth.text(40_000_000, 5), synth.longrange(30_000_000, 6)])
    one = codec.compress(x, 24, device=0)
    two, st = codec.compress(x, 24, devices=[0, 1], with_stats=True)
    assert two == one
    assert codec.decompress(two) == x.tobytes()
