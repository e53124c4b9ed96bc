"""Bit-exact parity at BASELINE.json's sizes: the engine's candidate lists (through the C ABI) against digests
of the REFERENCE's own finders (tests/golden/digests_big.json, made by tests/golden/make_golden_big.py from
oracle/_ref/libnlzm_ref.so in R2 mode), per finder and for all four together:

  C2            100 000 000 B text, -window:24 (ring shifts, position overflow into the check bits after 16 MiB)
  -window:28    140 MB long-range data and 140 MB drifting text (hist_bits stays 28: check field of 4 bits,
                rk_bits 21, bt_bits 17), 300 MB drifting text (P >= W for the last 32 MB: quirk 1 at 28 bits)

plus a brute-force window scan of sampled positions (nothing missing, nothing nearer) and engines on two
devices inside one process.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

GOLD = os.path.join(ROOT, "tests", "golden", "digests_big.json")
MASKS = {"ht2": 1, "ht3": 2, "bt4": 4, "rk256": 8, "all": 15}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a)).hexdigest()


class _Digest:
    """digest of a CSR that arrives block by block (same definition as make_golden_big.csr_digest)"""

    def __init__(self):
        self.c, self.d, self.l, self.steps = hashlib.sha256(), hashlib.sha256(), hashlib.sha256(), 0

    def add(self, off, raw):
        self.c.update(np.diff(off.astype(np.int64)).astype(np.uint8).tobytes())
        dist = raw["dist_lo"].astype(np.uint32) | ((raw["dist_hi"].astype(np.uint32) & 0x0FFF) << 16)
        self.d.update(dist.tobytes())
        self.l.update((raw["len"] & 0x1FF).astype(np.uint16).tobytes())
        self.steps += raw.size

    def hexdigest(self):
        return hashlib.sha256((self.c.hexdigest() + self.d.hexdigest() + self.l.hexdigest()).encode()).hexdigest()


def _engine_digest(lib, x, hb, mask, block):
    from nlzm_b200.matchfinder import MatchFinders
    from nlzm_b200 import sharding
    dg = _Digest()
    with MatchFinders(lib) as mf:
        mf.Init(hb, x, finder_mask=mask, max_range=block)
        for i, (b, e) in enumerate(sharding.split_blocks(0, x.size, block)):
            off, raw = mf.FindAndUpdate(b, e, slot=i & 1, copy=False)
            dg.add(off, raw)
    return dg


def _golden(key):
    if not os.path.exists(GOLD):
        pytest.fail("tests/golden/digests_big.json is missing")
    g = json.load(open(GOLD))
    if key not in g:
        pytest.skip(f"{key} not in digests_big.json")
    return g[key]


@pytest.mark.parametrize("finder", ["all", "bt4"])
def test_c2_full_size_bit_exact(cuda_lib, finder):
    from nlzm_b200 import synth
    want = _golden("text:100000000:24")
    x = synth.make("text", 100_000_000)
    assert hashlib.sha256(x.tobytes()).hexdigest() == want["input_sha256"]
    dg = _engine_digest(cuda_lib, x, 24, MASKS[finder], 100_000_000)
    assert dg.steps == want[finder]["steps"]
    assert dg.hexdigest() == want[finder]["sha256"]


@pytest.mark.parametrize("key", ["longrange:140000000:28", "text_drift:140000000:28", "text_drift:300000000:28"])
def test_window28_bit_exact_per_finder(cuda_lib, key):
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import geometry
    want = _golden(key)
    kind, n, hb = key.split(":")
    x = synth.make(kind, int(n))
    assert geometry(x.size, int(hb), cuda_lib).hist_bits == 28
    assert hashlib.sha256(x.tobytes()).hexdigest() == want["input_sha256"]
    for finder, mask in MASKS.items():
        dg = _engine_digest(cuda_lib, x, 28, mask, 1 << 27)       # blocks of 128 Mi positions: retained segments
        assert dg.steps == want[finder]["steps"], (key, finder)
        assert dg.hexdigest() == want[finder]["sha256"], (key, finder)


def _brute_bt4(x, a, W):
    """exhaustive BT4 at position a: for every length, the nearest earlier in-window position
    that shares at least that many bytes, as a staircase [(len, dist)] with strictly increasing len and dist"""
    n = x.size
    cap = min(264, n - a)
    if cap < 4:
        return []
    lo = max(0, a - (W - 1))
    # candidates share the first 4 bytes (hist_bits >= 19: the bucket index cannot collide on 2-3 bytes)
    v = x[a:a + 4]
    win = x[lo:a + 3]
    hit = np.flatnonzero((win[:-3] == v[0]) & (win[1:-2] == v[1]) & (win[2:-1] == v[2]) & (win[3:] == v[3])) + lo
    hit = hit[hit < a]
    steps, best = [], 3
    for q in hit[::-1]:                                  # nearest first
        if best >= cap:
            break
        d = a - int(q)
        m = 4
        while m < cap and x[q + m] == x[a + m]:
            m += 1
        mm = 2 + (d >= 256) + (d >= 4096) + (d >= (1 << 20))
        if m > best and m >= mm:
            steps.append((m, d))
            best = m
    # staircase keeps, per distance, the longest; nearer entries with smaller len stay
    return steps


def test_brute_force_window_scan_sampled(cuda_lib):
    """32 MB of text at -window:24, BT4 alone: for 900 sampled positions the engine's staircase equals a brute-force
    scan of the whole window (no candidate missing, none nearer with at least that length)"""
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.text(32_000_000, 51)
    W = 1 << 24
    with MatchFinders(cuda_lib) as mf:
        mf.Init(24, x, finder_mask=4)
        off, st = mf.FindAndUpdate()
    rng = np.random.default_rng(7)
    sample = np.concatenate([rng.integers(0, x.size, 700), rng.integers(W - 2000, W + 2000, 100),
                             np.arange(x.size - 100, x.size)])
    for a in sample:
        a = int(a)
        got = [(int(st["len"][j]), int(st["dist"][j])) for j in range(int(off[a]), int(off[a + 1]))]
        # the staircase lists steps by increasing len and dist; the brute force finds them nearest first
        want = _brute_bt4(x, a, W)
        assert got == want, (a, got, want)


def test_two_devices_in_one_process(cuda_lib, orc):
    """one host process driving engines on two GPUs (the deployment north_star describes): the shared-memory
    opt-in is per device; both engines must agree with the oracle, also through the segment hand-over"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from nlzm_b200 import synth, sharding
    from nlzm_b200.matchfinder import MatchFinders
    x = np.concatenate([synth.text(900_000, 11), synth.longrange(1_300_000, 12)])
    hb = 20
    ref = orc.find(x, hb, orc.F_ALL)
    with MatchFinders(cuda_lib) as m0, MatchFinders(cuda_lib) as m1:
        m0.Init(hb, x, device=0)
        m1.Init(hb, x, device=1)
        for mf in (m1, m0):                               # device 1 first: it must not inherit device 0's opt-in
            off, st = mf.FindAndUpdate()
            assert orc.csr_equal(ref, (off.astype(np.uint64), st["dist"], st["len"]))
        # position sharding with the peer copy of the neighbour's segments (NVLink, same process)
        cut = sharding.shard_range(x.size, 1, 2)[0]
        m0.drop_segments(); m1.drop_segments()
        m0.prepare(0, cut)
        m1.prepare(cut, x.size)
        for d in m0.export_segments():
            if d.pos_end > cut - ((1 << hb) - 1):
                m1.import_segment(d)
        o0, s0 = m0.FindAndUpdate(0, cut)
        o1, s1 = m1.FindAndUpdate(cut, x.size)
        assert m1.stats().segments_queried > 0
        got = sharding.concat_views([(0, cut, o0, s0), (cut, x.size, o1, s1)])
        assert orc.csr_equal(ref, got), orc.first_diff(ref, got)
