"""Drop-in check (north_star correctness checks 2 and 3): the reference's own parser, cost model and
rANS coder, unchanged, consuming the engine's candidates through the C++ host shim
(include/nlzm_mf_shim.hpp, binding = the sed patch in oracle/Makefile / INTEGRATION.md) must emit a
stream that the pristine reference decoder restores byte for byte, within 0.5 % of the reference's
own compressed size at the same -window."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _roundtrip(tmp_path, x, hb, emu, block_len):
    from oracle import refbind as rb
    src, ours, ref0, back = (str(tmp_path / n) for n in ("in.bin", "ours.nlzm", "r0.nlzm", "back.bin"))
    for f in (ours, ref0, back):
        if os.path.exists(f):
            os.remove(f)
    x.tofile(src)
    secs, served = rb.engine_fed_encode(src, ours, hb, emu=emu, block_len=block_len)
    rb.r0_cli(f"-window:{hb}", "c", src, ref0)              # pristine reference, same settings
    rb.r0_cli("d", ours, back)                              # pristine reference decoder
    assert open(back, "rb").read() == x.tobytes(), "engine-fed stream does not decode with the reference decoder"
    s_ours, s_ref = os.path.getsize(ours), os.path.getsize(ref0)
    assert abs(s_ours - s_ref) <= 0.005 * s_ref, (s_ours, s_ref)
    assert served > 0
    return s_ours, s_ref, secs


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} not built (make -C oracle gpu, needs /root/reference)")


@pytest.mark.parametrize("kind,n,hb", [("text", 150_000, 24), ("longrange", 200_000, 15), ("mixed", 120_000, 20)])
def test_engine_fed_reference_encoder_emulated(tmp_path, emu_lib, kind, n, hb):
    from oracle import refbind as rb
    from nlzm_b200 import synth
    _need(rb.REF_EMU_SO)
    _need(rb.REF_R0)
    _roundtrip(tmp_path, synth.make(kind, n), hb, emu=True, block_len=60_000)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,hb", [("text", 6_000_000, 24), ("longrange", 8_000_000, 24), ("mixed", 4_000_000, 22),
                                       ("text", 3_000_000, 15)])
def test_engine_fed_reference_encoder_gpu(tmp_path, cuda_lib, kind, n, hb):
    from oracle import refbind as rb
    from nlzm_b200 import synth
    _need(rb.REF_GPU_SO)
    _need(rb.REF_R0)
    s_ours, s_ref, secs = _roundtrip(tmp_path, synth.make(kind, n), hb, emu=False, block_len=1 << 21)
    print(f"{kind} {n} -window:{hb}: engine-fed {s_ours} B vs reference {s_ref} B ({(s_ours - s_ref) / s_ref * 100:+.3f} %), {secs:.1f} s")
