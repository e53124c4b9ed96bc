"""The order-independent algorithms (tests/algo_model.py) reproduce the oracle on small inputs."""
import numpy as np
import pytest

import algo_model as am


def _csr(orc, n, cands):
    if not cands:
        return orc.records_to_csr(n, [], [], [])
    a = np.array(cands)
    return orc.records_to_csr(n, a[:, 0], a[:, 1], a[:, 2])


CASES = [("text", 24_000, 15, 1), ("longrange", 80_000, 15, 2), ("mixed", 30_000, 15, 3)]


@pytest.mark.parametrize("kind,n,hb,seed", CASES)
def test_models(orc, kind, n, hb, seed):
    from nlzm_b200 import synth
    x = synth.make(kind, n, seed)
    xl = x.tolist()
    assert orc.csr_equal(_csr(orc, n, am.ht_model(xl, hb, 2)), orc.find(x, hb, orc.F_HT2))
    assert orc.csr_equal(_csr(orc, n, am.ht_model(xl, hb, 3)), orc.find(x, hb, orc.F_HT3))
    assert orc.csr_equal(_csr(orc, n, am.rk_model(xl, hb)), orc.find(x, hb, orc.F_RK256))
    if n <= 30_000:
        bt = am.bt4_model(xl, hb) + am.bt4_short_model(xl, hb)
        assert orc.csr_equal(_csr(orc, n, bt), orc.find(x, hb, orc.F_BT4))


def test_geometry_formula():
    """closed-form ring-shift epoch == simulation of encode_file's chunk loop"""
    for flen, hb in [(100_000_000, 24), (100_000_000, 15), (5_000_000, 17), (1_000_000_000, 28), (700_000, 15)]:
        g = am.Geometry(flen, hb)
        for k in range(0, (flen + g.cs - 1) // g.cs, max(1, flen // g.cs // 2000)):
            q = (k * g.cs) >> g.hb
            assert (q - 1 if q > 1 else 0) == g.epoch(k * g.cs)
