"""The C-ABI library loads and exports every symbol include/nlzm_mf.h declares; without a GPU it
refuses to create an engine (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "nlzm_mf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nlzm_mf_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    from nlzm_b200 import _lib, build
    build.build_cuda()
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in nlzm_mf.h but not exported"
    assert set(_lib.EXPORTS) == set(names)


def test_geometry_matches_oracle(orc):
    from nlzm_b200 import _lib
    from nlzm_b200.matchfinder import geometry
    L = _lib.load()
    for flen, hb in [(100_000_000, 24), (1_000_000_000, 28), (5_000_000, 17), (700_000, 15), (3000, 22),
                     (268_435_456, 28), (12, 15)]:
        g, o = geometry(flen, hb, L), orc.geometry(flen, hb)
        for f, _ in g._fields_:
            assert getattr(g, f) == getattr(o, f), (flen, hb, f)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nlzm_b200.matchfinder import MatchFinders, MatchFinderError
    import numpy as np
    with pytest.raises(MatchFinderError, match="no CPU fallback"):
        MatchFinders().Init(20, np.zeros(1000, np.uint8))
