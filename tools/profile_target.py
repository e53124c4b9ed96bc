"""Target for ncu: one find over the bench workload (or a smaller one) — no timing claims."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlzm_b200 import synth
from nlzm_b200.matchfinder import MatchFinders
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
hb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
kind = sys.argv[3] if len(sys.argv) > 3 else "text"
x = synth.make(kind, n)
with MatchFinders() as mf:
    mf.Init(hb, x)
    v = mf.find_device()
    print("steps", v.n_steps)
