"""Target for compute-sanitizer: several small finds (whole file and block mode, small and large windows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nlzm_b200 import synth
from nlzm_b200.matchfinder import MatchFinders
for kind, n, hb in [("text", 300_000, 15), ("longrange", 400_000, 16), ("zeros", 100_000, 15), ("mixed", 200_000, 20),
                    ("text", 1_500_000, 24), ("text", 1000, 15), ("text", 5, 15)]:
    x = synth.make(kind, n)
    with MatchFinders() as mf:
        mf.Init(hb, x)
        off, st = mf.FindAndUpdate()
        o2, s2 = mf.FindAndUpdate(n // 3, n, slot=1)
        print(kind, n, hb, st.size, s2.size, flush=True)
