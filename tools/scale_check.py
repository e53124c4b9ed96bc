"""Large-config check on the GPU box: run the engine over a config in blocks, verify size-independent
properties of every step (true match, strictly increasing, window, minimum length) and time it."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlzm_b200 import synth, sharding
from nlzm_b200.matchfinder import MatchFinders, unpack_steps

kind, n, hb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
block = int(sys.argv[4]) if len(sys.argv) > 4 else (1 << 28 if hb >= 27 else 1 << 27)   # big windows: big blocks (halo is re-merged per block)
t = time.time(); x = synth.make(kind, n); print(f"{kind} n={n} hb={hb}: generated in {time.time()-t:.1f}s", flush=True)
W = 1 << hb
with MatchFinders() as mf:
    mf.Init(hb, x, max_range=block)
    tot_ms = 0.0; tot_steps = 0
    for i, (b, e) in enumerate(sharding.split_blocks(0, n, block)):
        t = time.time(); off, st = mf.FindAndUpdate(b, e, slot=i & 1, copy=False); wall = time.time() - t; st = unpack_steps(st)
        s = mf.stats(); tot_ms += s.ms_total; tot_steps += st.size
        dist, ln = st["dist"].astype(np.int64), st["len"].astype(np.int64)
        pos = np.repeat(np.arange(b, e, dtype=np.int64), np.diff(off.astype(np.int64)))
        assert dist.size == 0 or (dist.min() >= 1 and dist.max() <= W - 1), "distance outside the window"
        assert (pos - dist >= 0).all()
        same = pos[1:] == pos[:-1]
        assert (ln[1:][same] > ln[:-1][same]).all() and (dist[1:][same] > dist[:-1][same]).all(), "steps not strictly increasing"
        mm = 2 + (dist >= 256) + (dist >= 4096) + (dist >= (1 << 20))
        assert (ln >= mm).all() and (ln <= 264).all(), "length bounds"
        assert (pos + ln <= n).all()
        assert (x[pos + ln - 1] == x[pos - dist + ln - 1]).all() and (x[pos] == x[pos - dist]).all(), "not a match"
        rng = np.random.default_rng(i)
        for j in rng.integers(0, max(pos.size, 1), 3000):
            if pos.size == 0: break
            a, d, l = int(pos[j]), int(dist[j]), int(ln[j])
            assert np.array_equal(x[a:a + l], x[a - d:a - d + l]), (a, d, l)
        print(f"  block [{b},{e}): {s.ms_total:.1f} ms device ({(e-b)/s.ms_total/1e3:.1f} MB/s), wall {wall*1e3:.0f} ms, steps {st.size} "
              f"({st.size/(e-b):.2f}/pos) rank={s.ms_rank:.0f} levels={s.ms_levels:.0f} ht={s.ms_ht:.0f} rk={s.ms_rk:.0f} merge={s.ms_merge:.0f} d2h={s.ms_d2h:.0f}", flush=True)
print(f"{kind} n={n} hb={hb}: OK, {tot_steps} steps, device {tot_ms:.1f} ms -> {n/tot_ms/1e3:.1f} MB/s")
