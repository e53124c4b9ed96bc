# scratch: first timing of the engine on the GPU box
import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlzm_b200 import synth, _lib
if os.environ.get('NLZM_MF_LIB'): _lib.LIB_PATH = os.environ['NLZM_MF_LIB']   # tuning variants (tools only)
from nlzm_b200.matchfinder import MatchFinders, profile, kernel_times
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
hb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
kind = sys.argv[3] if len(sys.argv) > 3 else 'text'
x = synth.make(kind, n)
mf = MatchFinders(); mf.Init(hb, x)
rb = int(os.environ.get('NLZM_BEGIN', '0')); re_ = int(os.environ.get('NLZM_END', str(n)))
for it in range(3):
    if it == 2: profile(True)
    t = time.time(); v = mf.find_device(rb, re_); dt = time.time() - t
    s = mf.stats()
    print(f'{kind} n={n} hb={hb} iter{it}: {dt*1e3:.1f} ms  {n/dt/1e6:.1f} MB/s steps={v.n_steps} tuples={s.tuples_last} '
          f'rank={s.ms_rank:.1f} levels={s.ms_levels:.1f} ht={s.ms_ht:.1f} rk={s.ms_rk:.1f} merge={s.ms_merge:.1f} total={s.ms_total:.1f}')
kt = kernel_times()
for k, (c, ms) in sorted(kt.items(), key=lambda kv: -kv[1][1]): print(f'  {k:28s} {c:5d} launches {ms:9.2f} ms')
t = time.time(); off, st = mf.FindAndUpdate(rb, re_); print('with D2H', time.time() - t, mf.stats().ms_d2h)
