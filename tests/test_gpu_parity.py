"""Parity of the CUDA engine (through the C ABI) with the CPU oracle — bit-exact candidate lists."""
import numpy as np
import pytest

from conftest import csr_from_find

pytestmark = pytest.mark.gpu

KINDS = [("text", 2_000_000, 24), ("text", 900_000, 15), ("text_drift", 1_500_000, 20),
         ("longrange", 3_000_000, 24), ("longrange", 1_200_000, 16), ("mixed", 1_000_000, 24),
         ("mixed", 600_000, 17), ("random", 500_000, 22), ("zeros", 300_000, 18)]


def _run(lib, x, hb, mask=15, cuts=None):
    from nlzm_b200.matchfinder import MatchFinders
    with MatchFinders(lib) as mf:
        mf.Init(hb, x, finder_mask=mask)
        if cuts is None:
            return csr_from_find(*mf.FindAndUpdate())
        offs, ds, ls, base = [np.zeros(1, np.uint64)], [], [], 0
        for i, (b, e) in enumerate(zip(cuts[:-1], cuts[1:])):
            off, st = mf.FindAndUpdate(b, e, slot=i & 1)
            offs.append(off[1:].astype(np.uint64) + base)
            base += int(off[-1])
            ds.append(st["dist"].copy())
            ls.append(st["len"].copy())
        return np.concatenate(offs), np.concatenate(ds), np.concatenate(ls)


@pytest.mark.parametrize("kind,n,hb", KINDS)
def test_all_finders_match_oracle(cuda_lib, orc, kind, n, hb):
    from nlzm_b200 import synth
    x = synth.make(kind, n)
    ref = orc.find(x, hb, orc.F_ALL)
    got = _run(cuda_lib, x, hb)
    assert orc.csr_equal(ref, got), orc.first_diff(ref, got)


@pytest.mark.parametrize("mask", [1, 2, 4, 8])
def test_each_finder_alone(cuda_lib, orc, mask):
    from nlzm_b200 import synth
    x = np.concatenate([synth.text(500_000, 21), synth.longrange(700_000, 22), synth.mixed(300_000, 23)])
    for hb in (16, 24):
        ref = orc.find(x, hb, mask)
        got = _run(cuda_lib, x, hb, mask)
        assert orc.csr_equal(ref, got), (hb, orc.first_diff(ref, got))


def test_block_mode_equals_whole_file(cuda_lib, orc):
    from nlzm_b200 import synth
    x = synth.longrange(1_500_000, 31)
    for hb in (15, 24):
        ref = orc.find(x, hb, orc.F_ALL)
        n = x.size
        for cuts in ([0, n // 3, 2 * n // 3 + 17, n], [0, 1000, n - 5, n]):
            got = _run(cuda_lib, x, hb, 15, cuts)
            assert orc.csr_equal(ref, got), (hb, cuts, orc.first_diff(ref, got))


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 255, 256, 257, 300, 1000, 14848, 14849])
def test_tiny_inputs(cuda_lib, orc, n):
    from nlzm_b200 import synth
    x = synth.text(max(n, 1), 3)[:n]
    ref = orc.find(x, 15, orc.F_ALL) if n else (np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint16))
    got = _run(cuda_lib, x, 15)
    assert orc.csr_equal(ref, got)


def test_ht_far_prefix_tables(cuda_lib, orc):
    """later ranges with a tiny HT margin: chains fall back to the coarse prefix tables (exact)"""
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    x = np.concatenate([synth.mixed(500_000, 33), synth.text(500_000, 34)])
    for hb in (15, 24):
        ref = orc.find(x, hb, 3)
        with MatchFinders(cuda_lib) as mf:
            mf.Init(hb, x, finder_mask=3)
            mf.set_option("ht_margin", 0)
            mf.set_option("ht_coarse_log", 12)
            offs, ds, ls, base = [np.zeros(1, np.uint64)], [], [], 0
            cuts = [0, 300_000, 700_001, x.size]
            for i, (b, e) in enumerate(zip(cuts[:-1], cuts[1:])):
                off, st = mf.FindAndUpdate(b, e, slot=i & 1)
                offs.append(off[1:].astype(np.uint64) + base)
                base += int(off[-1])
                ds.append(st["dist"].copy())
                ls.append(st["len"].copy())
        got = (np.concatenate(offs), np.concatenate(ds), np.concatenate(ls))
        assert orc.csr_equal(ref, got), (hb, orc.first_diff(ref, got))


def test_submit_fetch_pipeline(cuda_lib, orc):
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.text(1_000_000, 41)
    ref = orc.find(x, 24, orc.F_ALL)
    n = x.size
    cuts = [0, n // 4, n // 2, 3 * n // 4, n]
    with MatchFinders(cuda_lib) as mf:
        mf.Init(24, x)
        offs, ds, ls, base = [np.zeros(1, np.uint64)], [], [], 0
        mf.submit(cuts[0], cuts[1], 0)
        for i in range(4):
            if i + 1 < 4:
                mf.submit(cuts[i + 1], cuts[i + 2], (i + 1) & 1)
            off, st = mf.fetch(i & 1)
            offs.append(off[1:].astype(np.uint64) + base)
            base += int(off[-1])
            ds.append(st["dist"])
            ls.append(st["len"])
    got = (np.concatenate(offs), np.concatenate(ds), np.concatenate(ls))
    assert orc.csr_equal(ref, got)


def test_properties_at_scale(cuda_lib):
    """32 MB text, -window:24: every step is a true, maximal match; steps strictly increase; distances
    stay inside the window (validity only: completeness is tests/test_gpu_scale.py's job)."""
    from nlzm_b200 import synth
    x = synth.text(32_000_000, 51)
    off, st = _run(cuda_lib, x, 24)[0], None
    from nlzm_b200.matchfinder import MatchFinders
    with MatchFinders(cuda_lib) as mf:
        mf.Init(24, x)
        off, st = mf.FindAndUpdate()
    dist, ln = st["dist"].astype(np.int64), st["len"].astype(np.int64)
    n = x.size
    pos = np.repeat(np.arange(n, dtype=np.int64), np.diff(off.astype(np.int64)))
    assert dist.min() >= 1 and dist.max() <= (1 << 24) - 1
    assert (pos - dist >= 0).all()
    same = pos[1:] == pos[:-1]
    assert (ln[1:][same] > ln[:-1][same]).all() and (dist[1:][same] > dist[:-1][same]).all()
    mm = 2 + (dist >= 256) + (dist >= 4096) + (dist >= (1 << 20))
    assert (ln >= mm).all() and (ln <= 264).all()
    # true matches: compare the last byte and a few random inner bytes of every step, all bytes of a sample
    assert (x[pos + ln - 1] == x[pos - dist + ln - 1]).all()
    assert (x[pos] == x[pos - dist]).all()
    rng = np.random.default_rng(0)
    for j in rng.integers(0, pos.size, 20000):
        a, d, l = int(pos[j]), int(dist[j]), int(ln[j])
        assert np.array_equal(x[a:a + l], x[a - d:a - d + l])
    # completeness (no candidate missing, none nearer) is checked against a brute-force scan of the whole window in
    # tests/test_gpu_scale.py::test_brute_force_window_scan_sampled and bit-exactly at full size in the digest tests


def _check_properties(x, off, st, begin, W):
    dist, ln = st["dist"].astype(np.int64), st["len"].astype(np.int64)
    pos = np.repeat(np.arange(begin, begin + off.size - 1, dtype=np.int64), np.diff(off.astype(np.int64)))
    assert dist.size == 0 or (dist.min() >= 1 and dist.max() <= W - 1)
    assert (pos - dist >= 0).all() and (pos + ln <= x.size).all()
    same = pos[1:] == pos[:-1]
    assert (ln[1:][same] > ln[:-1][same]).all() and (dist[1:][same] > dist[:-1][same]).all()
    mm = 2 + (dist >= 256) + (dist >= 4096) + (dist >= (1 << 20))
    assert (ln >= mm).all() and (ln <= 264).all()
    assert (x[pos] == x[pos - dist]).all() and (x[pos + ln - 1] == x[pos - dist + ln - 1]).all()
    rng = np.random.default_rng(1)
    for j in rng.integers(0, max(pos.size, 1), 5000):
        a, d, l = int(pos[j]), int(dist[j]), int(ln[j])
        assert np.array_equal(x[a:a + l], x[a - d:a - d + l])


def test_properties_256mb_window(cuda_lib):
    """-window:28 needs >= 128 MiB of input (the reference shrinks the window to the file): 140 MB of
    long-range data in two blocks; every step must be a true in-window match, strictly increasing."""
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders, geometry, unpack_steps
    x = synth.longrange(140_000_000, 77)
    assert geometry(x.size, 28, cuda_lib).hist_bits == 28
    with MatchFinders(cuda_lib) as mf:
        mf.Init(28, x, max_range=1 << 27)
        total = 0
        for i, (b, e) in enumerate([(0, 1 << 26), (1 << 26, x.size)]):
            off, st = mf.FindAndUpdate(b, e, slot=i, copy=False)
            st = unpack_steps(st)
            _check_properties(x, off, st, b, 1 << 28)
            total += st.size
    assert total > x.size          # redundant data: several steps per position


@pytest.mark.parametrize("kind", ["zeros_ones", "period", "ab", "words", "text"])
def test_retained_segments_adversarial_and_overflow(cuda_lib, orc, kind):
    """consecutive blocks (retained segments, text-order cross merge), small HT coarse tiles (cell snapshot + warp
    scans of the far prefix) and a forced candidate-buffer overflow, on inputs whose suffixes tie for hundreds of
    bytes and end in a zero run"""
    from nlzm_b200 import synth
    from test_fuzz import _gen
    from nlzm_b200.matchfinder import MatchFinders
    n = 700_000
    x = synth.text(n, 5) if kind == "text" else _gen(kind, n, np.random.default_rng(11))
    if kind == "zeros_ones":
        x[-3000:] = 0
    hb = 17
    ref = orc.find(x, hb, orc.F_ALL)
    cuts = [0, 150_000, 290_001, 431_000, 650_000, n]
    with MatchFinders(cuda_lib) as mf:
        mf.Init(hb, x)
        mf.set_option("ht_coarse_log", 13)
        if kind in ("text", "words"):
            mf.set_option("tuple_cap_extra", 0)
            mf.set_option("tuple_cap_mult", 1)
        offs, ds, ls, base, used = [np.zeros(1, np.uint64)], [], [], 0, []
        for i, (b, e) in enumerate(zip(cuts[:-1], cuts[1:])):
            off, st = mf.FindAndUpdate(b, e, slot=i & 1)
            used.append(int(mf.stats().segments_queried))
            offs.append(off[1:].astype(np.uint64) + base)
            base += int(off[-1])
            ds.append(st["dist"].copy())
            ls.append(st["len"].copy())
    got = (np.concatenate(offs), np.concatenate(ds), np.concatenate(ls))
    assert orc.csr_equal(ref, got), (kind, orc.first_diff(ref, got))
    assert all(u > 0 for u in used[1:]), used
