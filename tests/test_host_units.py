"""Building blocks of the host pipeline (nlzm_b200/csrc/host/*.hpp), checked in C++ without the engine:
staircase merge == per-step updates, frame writer/reader round trip on adapting tables, the integer
price list, distance slot codes. The program is tests/host_units.cpp."""
import os
import subprocess

from conftest import ROOT


def test_host_units(tmp_path):
    exe = str(tmp_path / "host_units")
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "host_units.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "host units ok" in out.stdout, out.stdout + out.stderr


def test_decoder_under_sanitizers(tmp_path, emu_lib):
    """tests/decoder_sanitize.cpp: the stream reader built with -fsanitize=address,undefined over a
    reference-written stream, its truncations and 400 bit flips."""
    import numpy as np
    from nlzm_b200 import synth
    emu = os.path.join(ROOT, "tests", "emu")
    exe = str(tmp_path / "decoder_sanitize")
    probe = subprocess.run(["g++", "-fsanitize=address,undefined", "-x", "c++", "-", "-o", str(tmp_path / "probe")],
                           input="int main(){return 0;}", text=True, capture_output=True)
    if probe.returncode != 0:
        import pytest
        pytest.skip("no sanitizer runtime for g++ here")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           os.path.join(ROOT, "tests", "decoder_sanitize.cpp"),
                           os.path.join(ROOT, "nlzm_b200", "csrc", "host", "codec.cpp"), "-o", exe,
                           "-L" + emu, "-lnlzm_mf_emu", "-Wl,-rpath," + emu])
    want = str(tmp_path / "want.bin")
    synth.make("mixed", 40_000).tofile(want)
    stream = os.path.join(ROOT, "tests", "golden", "streams", "r0_mixed_40000_w20.nlzm")
    out = subprocess.run([exe, stream, want], capture_output=True, text=True)
    assert out.returncode == 0 and "decoder sanitize ok" in out.stdout, out.stdout + out.stderr[-2000:]
