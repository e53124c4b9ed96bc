"""Pins the CPU oracle (oracle/nlzm_oracle.c): against the golden fixtures generated from the
reference itself, and — when the reference library was built here — against the reference live."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
BIT = {0: 1, 1: 2, 2: 4, 3: 8}


def _digest(csr):
    h = hashlib.sha256()
    for a in csr:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "small_*.npz"))))
def test_oracle_matches_golden_records(orc, path):
    z = np.load(path)
    x, hb = z["x"], int(z["hist_bits"])
    for f, bit in BIT.items():
        sel = z["finder"] == f
        ref = orc.records_to_csr(x.size, z["pos"][sel], z["dist"][sel], z["len"][sel])
        got = orc.find(x, hb, bit)
        assert orc.csr_equal(ref, got), (os.path.basename(path), f, orc.first_diff(ref, got))
    ref = orc.records_to_csr(x.size, z["pos"], z["dist"], z["len"])
    assert orc.csr_equal(ref, orc.find(x, hb, orc.F_ALL))


def test_oracle_matches_golden_digests(orc):
    from nlzm_b200 import synth
    dig = json.load(open(os.path.join(GOLD, "digests.json")))
    assert len(dig) >= 9
    for key, want in dig.items():
        kind, n, hb = key.split(":")
        x = synth.make(kind, int(n))
        assert hashlib.sha256(x.tobytes()).hexdigest() == want["input_sha256"], f"synth.{kind} is not reproducible"
        got = orc.find(x, int(hb), orc.F_ALL)
        got = (got[0].astype(np.uint64), got[1].astype(np.uint32), got[2].astype(np.uint16))
        assert got[1].size == want["steps"], key
        assert _digest(got) == want["sha256"], key


def test_oracle_vs_reference_live(orc):
    from oracle import refbind as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libnlzm_ref.so not built (needs /root/reference)")
    from nlzm_b200 import synth
    for kind, n, hb, seed in [("text", 300_000, 16, 61), ("longrange", 500_000, 24, 62), ("mixed", 250_000, 15, 63)]:
        x = synth.make(kind, n, seed)
        recs, _ = rb.matchfind(x, hb, mode=rb.R2)
        for f, bit in BIT.items():
            sel = recs["finder"] == f
            ref = orc.records_to_csr(x.size, recs["pos"][sel], recs["dist"][sel], recs["len"][sel])
            got = orc.find(x, hb, bit)
            assert orc.csr_equal(ref, got), (kind, f, orc.first_diff(ref, got))


def test_reference_driver_equals_real_encoder(tmp_path):
    """The matcher-only driver around the reference's finder objects reports exactly what the hooked
    real encoder (encode_file -> parse_table) reports in R2 mode, and the stream round-trips."""
    from oracle import refbind as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libnlzm_ref.so not built (needs /root/reference)")
    from nlzm_b200 import synth
    x = synth.longrange(400_000, 71)
    src, dst, back = str(tmp_path / "in"), str(tmp_path / "out.nlzm"), str(tmp_path / "back")
    x.tofile(src)
    for hb in (15, 24):
        a = rb.encode_dump(src, dst, hb, mode=rb.R2)
        b, _ = rb.matchfind(x, hb, mode=rb.R2)
        assert np.array_equal(a, b)
        rb.decode(dst, back)
        assert open(back, "rb").read() == x.tobytes()


def test_oracle_bt_cap_is_honoured(orc):
    """bt_max_tests=256 (as shipped) may only remove BT4 steps, never invent one."""
    from nlzm_b200 import synth
    x = synth.mixed(300_000, 81)
    full = orc.find(x, 17, orc.F_BT4, 0)
    capped = orc.find(x, 17, orc.F_BT4, 256)
    assert capped[1].size <= full[1].size
