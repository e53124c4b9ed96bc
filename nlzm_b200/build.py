"""Builds the libraries in-tree: nlzm_b200/csrc/libnlzm_mf.so (CUDA engine, sm_100a only) and
nlzm_b200/csrc/libnlzm_codec.so (C++ host pipeline on top of it)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libnlzm_mf.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    srcs = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, os.path.join(CSRC, "engine.cu")]
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


CODEC_OUT = os.path.join(CSRC, "libnlzm_codec.so")


def build_codec(force: bool = False) -> str:
    """Host pipeline (parser + stream writer / reader); links the engine library next to it."""
    build_cuda()
    srcs = glob.glob(os.path.join(CSRC, "host", "*")) + glob.glob(os.path.join(HERE, "..", "include", "*"))
    if not force and os.path.exists(CODEC_OUT) and \
            all(os.path.getmtime(s) <= os.path.getmtime(CODEC_OUT) for s in srcs + [OUT]):
        return CODEC_OUT
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-O3", "-g", "-std=c++17", "-Wall", "-fPIC", "-shared", os.path.join(CSRC, "host", "codec.cpp"),
           "-o", CODEC_OUT, "-L" + CSRC, "-lnlzm_mf", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd, cwd=CSRC)
    return CODEC_OUT
