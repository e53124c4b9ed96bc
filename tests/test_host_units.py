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
