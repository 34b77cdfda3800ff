"""Small single-GPU driver for ncu: a few waves of the C2 sampler (South 1NT)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridge_mc_b200 import bidding
from bridge_mc_b200.engine import Constraints, Engine

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 26
eng = Engine(0)
if cfg == "c1":
    cons = None
elif cfg == "c3":
    p = bidding.openings.passes()
    cons = Constraints(None, [f"(and {p} T)", p, bidding.openings.chooses((1, "NT")), p])
elif cfg == "c5":
    cons = Constraints(None, [bidding.nt_responses(2).chooses((6, "NT")), None, bidding.openings.chooses((2, "NT")), None])
else:
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
if len(sys.argv) > 3 and cons is not None:
    cons.set_mode(int(sys.argv[3]))
for i in range(3):
    rec, tal, st = eng.sample(cons, n_raw=n, first_index=i << 40, cap=0)
    print(cfg, "raw", tal["raw"], "acc", tal["accepted"], "ms", round(st.kernel_ms, 3),
          "deals/s %.3e" % (tal["raw"] / st.kernel_ms * 1e3), "mode", cons.mode if cons else 1)
