"""Small adversarial inputs through the emulated engine vs the oracle: tiny alphabets, periodic data,
long runs, near-EOF caps, block cuts at odd places."""
import numpy as np
import pytest

from conftest import csr_from_find


def _gen(kind, n, rng):
    if kind == "ab":
        return rng.integers(97, 99, n, dtype=np.uint8)
    if kind == "abc_runs":
        x = np.repeat(rng.integers(97, 100, n // 3 + 1, dtype=np.uint8), rng.integers(1, 9, n // 3 + 1))[:n]
        return np.ascontiguousarray(x)
    if kind == "period":
        p = int(rng.integers(1, 40))
        base = rng.integers(0, 256, p, dtype=np.uint8)
        x = np.tile(base, n // p + 1)[:n].copy()
        flips = rng.integers(0, n, max(1, n // 300))
        x[flips] ^= 1
        return x
    if kind == "zeros_ones":
        x = np.zeros(n, np.uint8)
        x[rng.integers(0, n, max(1, n // 500))] = 1
        return x
    if kind == "words":
        vocab = [bytes(rng.integers(97, 123, int(rng.integers(1, 7)), dtype=np.uint8)) for _ in range(12)]
        out = b" ".join(vocab[int(i)] for i in rng.integers(0, 12, n // 3 + 1))
        return np.frombuffer(out[:n].ljust(n, b"."), dtype=np.uint8).copy()
    raise ValueError(kind)


KINDS = ["ab", "abc_runs", "period", "zeros_ones", "words"]


@pytest.mark.parametrize("kind", KINDS)
def test_emu_fuzz(emu_lib, orc, kind):
    _fuzz(emu_lib, orc, kind, (7, 263, 264, 265, 700, 4099, 9000, 40_000))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_fuzz(cuda_lib, orc, kind):
    _fuzz(cuda_lib, orc, kind, (7, 263, 264, 265, 700, 4099, 9000, 40_000, 300_000))


def _fuzz(lib, orc, kind, sizes):
    from nlzm_b200.matchfinder import MatchFinders
    emu_lib = lib
    rng = np.random.default_rng(sum(kind.encode()))
    for n in sizes:
        x = _gen(kind, n, rng)
        ref = orc.find(x, 15, orc.F_ALL)
        with MatchFinders(emu_lib) as mf:
            mf.Init(15, x)
            got = csr_from_find(*mf.FindAndUpdate())
            assert orc.csr_equal(ref, got), (kind, n, orc.first_diff(ref, got))
            if n > 600:
                cut = int(rng.integers(1, n - 1))
                o1, s1 = mf.FindAndUpdate(0, cut, slot=0)
                o2, s2 = mf.FindAndUpdate(cut, n, slot=1)
                off = np.concatenate([o1.astype(np.uint64), o2[1:].astype(np.uint64) + int(o1[-1])])
                got2 = (off, np.concatenate([s1["dist"], s2["dist"]]), np.concatenate([s1["len"], s2["len"]]))
                assert orc.csr_equal(ref, got2), (kind, n, cut, orc.first_diff(ref, got2))
