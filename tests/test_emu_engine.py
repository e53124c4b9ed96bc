"""Kernel bodies + host orchestration of the engine, run sequentially on the CPU (tests/emu), against
the oracle. This checks the logic that the GPU tests re-check on the real device."""
import numpy as np
import pytest

from conftest import csr_from_find


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


@pytest.mark.parametrize("kind,n,hb", [("text", 90_000, 24), ("text", 120_000, 15), ("longrange", 140_000, 24),
                                       ("longrange", 110_000, 16), ("mixed", 100_000, 16), ("zeros", 30_000, 15),
                                       ("random", 50_000, 20)])
def test_emu_all_finders(emu_lib, orc, kind, n, hb):
    from nlzm_b200 import synth
    x = synth.make(kind, n)
    ref = orc.find(x, hb, orc.F_ALL)
    got = _run(emu_lib, x, hb)
    assert orc.csr_equal(ref, got), orc.first_diff(ref, got)


@pytest.mark.parametrize("mask", [1, 2, 4, 8])
def test_emu_each_finder(emu_lib, orc, mask):
    from nlzm_b200 import synth
    x = np.concatenate([synth.text(40_000, 21), synth.longrange(70_000, 22), synth.mixed(30_000, 23)])
    ref = orc.find(x, 15, mask)
    got = _run(emu_lib, x, 15, mask)
    assert orc.csr_equal(ref, got), orc.first_diff(ref, got)


def test_emu_block_mode(emu_lib, orc):
    from nlzm_b200 import synth
    x = synth.longrange(110_000, 31)
    n = x.size
    for hb in (15, 24):
        ref = orc.find(x, hb, orc.F_ALL)
        for cuts in ([0, n // 3, 2 * n // 3 + 17, n], [0, 1000, n - 5, n], [0, n - 2, n]):
            got = _run(emu_lib, x, hb, 15, cuts)
            assert orc.csr_equal(ref, got), (hb, cuts)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 255, 256, 257, 300, 1000])
def test_emu_tiny(emu_lib, orc, n):
    from nlzm_b200 import synth
    x = synth.text(max(n, 1), 3)[:n]
    ref = orc.find(x, 15, orc.F_ALL) if n else (np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint16))
    assert orc.csr_equal(ref, _run(emu_lib, x, 15))


def test_emu_ht_far_prefix(emu_lib, orc):
    """HT stage with a tiny margin: chains of later ranges must fall back to the coarse prefix tables"""
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.mixed(150_000, 33)
    for hb, mask in ((15, 3), (16, 3)):
        ref = orc.find(x, hb, mask)
        with MatchFinders(emu_lib) as mf:
            mf.Init(hb, x, finder_mask=mask)
            mf.set_option("ht_margin", 0)
            mf.set_option("ht_coarse_log", 11)
            offs, ds, ls, base = [np.zeros(1, np.uint64)], [], [], 0
            cuts = [0, 50_000, 100_001, x.size]
            for i, (b, e) in enumerate(zip(cuts[:-1], cuts[1:])):
                off, st = mf.FindAndUpdate(b, e, slot=i & 1)
                offs.append(off[1:].astype(np.uint64) + base)
                base += int(off[-1])
                ds.append(st["dist"].copy())
                ls.append(st["len"].copy())
        got = (np.concatenate(offs), np.concatenate(ds), np.concatenate(ls))
        assert orc.csr_equal(ref, got), (hb, orc.first_diff(ref, got))


def test_emu_golden(emu_lib, orc):
    """committed fixtures that came from the reference's real encoder"""
    import glob, os
    from conftest import ROOT
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "small_*.npz")))[:3]:
        z = np.load(path)
        ref = orc.records_to_csr(z["x"].size, z["pos"], z["dist"], z["len"])
        got = _run(emu_lib, z["x"], int(z["hist_bits"]))
        assert orc.csr_equal(ref, got), os.path.basename(path)


def test_error_paths(emu_lib):
    from nlzm_b200.matchfinder import MatchFinders, MatchFinderError
    mf = MatchFinders(emu_lib)
    mf.Init(15, np.zeros(100, np.uint8))
    with pytest.raises(MatchFinderError):
        mf.FindAndUpdate(50, 200)
    with pytest.raises(MatchFinderError):
        mf.FindAndUpdate(0, 10, slot=3)
    with pytest.raises(MatchFinderError):
        mf.fetch(0)
    mf.Release()


def _blocks(mf, cuts):
    offs, ds, ls, base, used = [np.zeros(1, np.uint64)], [], [], 0, []
    for i, (b, e) in enumerate(zip(cuts[:-1], cuts[1:])):
        off, st = mf.FindAndUpdate(b, e, slot=i & 1)
        used.append(int(mf.stats().segments_queried))
        offs.append(off[1:].astype(np.uint64) + base)
        base += int(off[-1])
        ds.append(st["dist"].copy())
        ls.append(st["len"].copy())
    return (np.concatenate(offs), np.concatenate(ds), np.concatenate(ls)), used


@pytest.mark.parametrize("kind,hb", [("text", 15), ("longrange", 15), ("mixed", 16), ("zeros", 15)])
def test_emu_retained_segments(emu_lib, orc, kind, hb):
    """consecutive ranges query the retained segments of the ranges before them instead of re-ranking the
    window; the result equals the oracle (and the from-scratch path, option retain=0)"""
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    n = 150_000 if kind != "zeros" else 80_000
    x = synth.make(kind, n)
    ref = orc.find(x, hb, orc.F_ALL)
    for cuts in ([0, 40_000, 47_000, n - 30_001, n], [0, 9_000, 18_000, 27_000, 36_000, 45_000, 54_000, n - 9_000, n]):
        with MatchFinders(emu_lib) as mf:
            mf.Init(hb, x)
            got, used = _blocks(mf, cuts)
            assert orc.csr_equal(ref, got), (cuts, orc.first_diff(ref, got))
            assert used[0] == 0 and all(u > 0 for u in used[1:]), used
        with MatchFinders(emu_lib) as mf:
            mf.Init(hb, x)
            mf.set_option("retain", 0)
            got, used = _blocks(mf, cuts)
            assert orc.csr_equal(ref, got) and not any(used)


def test_emu_retained_falls_back_when_not_consecutive(emu_lib, orc):
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.text(120_000, 5)
    ref = orc.find(x, 15, orc.F_ALL)
    with MatchFinders(emu_lib) as mf:
        mf.Init(15, x)
        mf.FindAndUpdate(0, 30_000)
        off, st = mf.FindAndUpdate(70_000, 120_000, slot=1)          # gap: window re-ranked with the range
        assert mf.stats().segments_queried == 0
        lo, hi = int(ref[0][70_000]), int(ref[0][120_000])
        assert np.array_equal(st["dist"], ref[1][lo:hi]) and np.array_equal(st["len"], ref[2][lo:hi])
        assert np.array_equal(off.astype(np.uint64), ref[0][70_000:] - ref[0][70_000])


@pytest.mark.parametrize("world", [2, 3, 5])
def test_emu_sharded_engines_import_segments(emu_lib, orc, world):
    """position sharding: every engine prepares its own range, imports the segments behind it from the
    engines that own them, then finds; the concatenation equals the oracle"""
    from nlzm_b200 import synth, sharding
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.longrange(160_000, 9)
    hb = 16                                               # W = 65536: a shard of 32000 needs two or three neighbours
    ref = orc.find(x, hb, orc.F_ALL)
    W = 1 << hb
    engines = [MatchFinders(emu_lib) for _ in range(world)]
    ranges = [sharding.shard_range(x.size, r, world) for r in range(world)]
    try:
        for mf, (b, e) in zip(engines, ranges):
            mf.Init(hb, x)
            mf.prepare(b, e)
        descs = [mf.export_segments() for mf in engines]
        parts = []
        for r, (mf, (b, e)) in enumerate(zip(engines, ranges)):
            for q in range(r):
                for d in descs[q]:
                    if d.pos_end > max(0, b - (W - 1)):
                        mf.import_segment(bytes(d))
            off, st = mf.FindAndUpdate(b, e)
            if r > 0:
                assert mf.stats().segments_queried > 0
            parts.append((b, e, off, st))
        got = sharding.concat_views(parts)
        assert orc.csr_equal(ref, got), orc.first_diff(ref, got)
    finally:
        for mf in engines:
            mf.Release()


def test_emu_sharded_multi_block_ranges(emu_lib, orc):
    """a shard larger than one engine call: its first block is prepared, the later blocks are found (they query the
    blocks before them), the neighbour's segments arrive, and the first block finishes last"""
    from nlzm_b200 import synth, sharding
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.text(150_000, 19)
    hb = 15
    W = 1 << hb
    ref = orc.find(x, hb, orc.F_ALL)
    world = 2
    engines = [MatchFinders(emu_lib) for _ in range(world)]
    ranges = [sharding.shard_range(x.size, r, world) for r in range(world)]
    blocks = [sharding.split_blocks(b, e, 34_000) for (b, e) in ranges]      # blocks of at least one window
    parts = []
    try:
        for mf, bl in zip(engines, blocks):
            mf.Init(hb, x)
            mf.prepare(*bl[0])
            for i, (b, e) in enumerate(bl[1:]):
                off, st = mf.FindAndUpdate(b, e, slot=i & 1)
                assert mf.stats().segments_queried > 0
                parts.append((b, e, off, st))
        # what the next shard cannot reach is cut off before the export (a straddling segment is compacted)
        engines[0].trim_segments(blocks[0][-1][1] - (W - 1))
        descs = [mf.export_segments() for mf in engines]
        assert descs[0][0].pos_begin == blocks[0][-1][1] - (W - 1)
        for r, (mf, bl) in enumerate(zip(engines, blocks)):
            b0 = bl[0][0]
            for q in range(r):
                for d in descs[q]:
                    if d.pos_end > max(0, b0 - (W - 1)) and d.pos_end <= b0:
                        mf.import_segment(bytes(d))
            off, st = mf.FindAndUpdate(*bl[0])
            parts.append((bl[0][0], bl[0][1], off, st))
        got = sharding.concat_views(parts)
        assert orc.csr_equal(ref, got), orc.first_diff(ref, got)
    finally:
        for mf in engines:
            mf.Release()


@pytest.mark.parametrize("kind", ["zeros_ones", "period", "ab", "abc_runs", "words"])
def test_emu_retained_segments_adversarial(emu_lib, orc, kind):
    """retained segments on inputs whose suffixes tie for hundreds of bytes and end in zero runs at the end of
    the file (the cross-segment merge order must agree with both rank orders there)"""
    from test_fuzz import _gen
    from nlzm_b200.matchfinder import MatchFinders
    rng = np.random.default_rng(sum(kind.encode()) + 7)
    for n, cuts in ((31_000, [0, 20_000, 31_000]), (50_000, [0, 9_000, 30_000, 49_700, 50_000])):
        x = _gen(kind, n, rng)
        if kind == "zeros_ones":
            x[-300:] = 0
        ref = orc.find(x, 15, orc.F_ALL)
        with MatchFinders(emu_lib) as mf:
            mf.Init(15, x)
            got, used = _blocks(mf, cuts)
        assert orc.csr_equal(ref, got), (kind, n, orc.first_diff(ref, got))
        assert all(u > 0 for u in used[1:])


@pytest.mark.parametrize("kind,hb,clog", [("text", 15, 10), ("mixed", 16, 11), ("zeros_ones", 15, 10), ("period", 15, 12),
                                          ("longrange", 16, 10)])
def test_emu_ht_cell_snapshot(emu_lib, orc, kind, hb, clog):
    """HT2/HT3 of later ranges: per-position data only from the coarse tile before the range on; chains that reach
    further back end in the snapshot of the cell contents at that point (ring shifts clear cell 0 in between)"""
    from nlzm_b200 import synth
    from test_fuzz import _gen
    from nlzm_b200.matchfinder import MatchFinders
    n = 140_000
    x = synth.make(kind, n) if kind in ("text", "mixed", "longrange") else _gen(kind, n, np.random.default_rng(5))
    for mask in (3, 15):
        ref = orc.find(x, hb, mask)
        with MatchFinders(emu_lib) as mf:
            mf.Init(hb, x, finder_mask=mask)
            mf.set_option("ht_coarse_log", clog)
            got, _ = _blocks(mf, [0, 33_000, 71_111, 100_000, n])
        assert orc.csr_equal(ref, got), (kind, mask, orc.first_diff(ref, got))


@pytest.mark.parametrize("world,n", [(2, 220_000), (5, 200_000)])
def test_emu_published_segments(emu_lib, orc, world, n):
    """export through the engine's export buffer (what other processes map once): copies cut down to the reach of the
    next shard, a straddling block compacted; two rounds on the same engines"""
    from nlzm_b200 import synth, sharding
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.text(n, 3)
    hb = 16
    W = 1 << hb
    ref = orc.find(x, hb, orc.F_ALL)
    engines = [MatchFinders(emu_lib) for _ in range(world)]
    ranges = [sharding.shard_range(x.size, r, world) for r in range(world)]
    try:
        for step in range(2):
            for mf, (b, e) in zip(engines, ranges):
                if step == 0:
                    mf.Init(hb, x)
                mf.prepare(b, e)
            descs = [mf.publish_segments(max(0, e - (W - 1))) for mf, (b, e) in zip(engines, ranges)]
            assert all(d.pos_begin >= max(0, e - (W - 1)) for ds, (b, e) in zip(descs, ranges) for d in ds)
            parts = []
            for r, (mf, (b, e)) in enumerate(zip(engines, ranges)):
                for q in range(r - 1, -1, -1):
                    for d in descs[q]:
                        if d.pos_end > max(0, b - (W - 1)) and d.pos_end <= b:
                            mf.import_segment(bytes(d))
                off, st = mf.FindAndUpdate(b, e)
                parts.append((b, e, off, st))
            assert orc.csr_equal(ref, sharding.concat_views(parts))
    finally:
        for mf in engines:
            mf.Release()


def test_emu_tuple_overflow_is_retried(emu_lib, orc):
    """a candidate buffer that is too small: the engine grows it and redoes the range (also a prepared one, and one
    that queries retained segments) — same result"""
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    x = synth.text(90_000, 77)
    ref = orc.find(x, 15, orc.F_ALL)
    with MatchFinders(emu_lib) as mf:
        mf.Init(15, x)
        mf.set_option("tuple_cap_extra", 0)
        mf.set_option("tuple_cap_mult", 1)
        got, used = _blocks(mf, [0, 40_000, 90_000])
        assert orc.csr_equal(ref, got) and used[1] > 0
    with MatchFinders(emu_lib) as mf:
        mf.Init(15, x)
        mf.set_option("tuple_cap_extra", 0)
        mf.set_option("tuple_cap_mult", 1)
        mf.prepare(0, 40_000)
        got, _ = _blocks(mf, [0, 40_000])
        assert orc.csr_equal((ref[0][:40_001], ref[1][:int(ref[0][40_000])], ref[2][:int(ref[0][40_000])]), got)


@pytest.mark.parametrize("kind", ["zeros", "longrange", "period"])
def test_emu_rk_restart_point(emu_lib, orc, kind):
    """stage R of a late range: hits are looked up from a little before the range only and the carried-match machine
    starts behind a stretch of 65536 hit-free positions (sparse hits), or the range is redone from the last ring shift
    (dense hits); both must equal the sequential reference"""
    from nlzm_b200 import synth
    from nlzm_b200.matchfinder import MatchFinders
    from test_fuzz import _gen
    n = 700_000
    x = _gen(kind, n, np.random.default_rng(3)) if kind == "period" else synth.make(kind, n)
    ref = orc.find(x, 24, orc.F_RK256)
    with MatchFinders(emu_lib) as mf:
        mf.Init(24, x, finder_mask=orc.F_RK256)
        mf.set_option("rk_restart", 70_000)
        got, _ = _blocks(mf, [0, 260_000, 480_001, n])
    assert orc.csr_equal(ref, got), (kind, orc.first_diff(ref, got))
    assert ref[1].size > 0
