"""Small single-GPU driver for ncu: a few launches of one sampler configuration.
    python tools/prof_c2.py [c1|c2|c3|c4|c5|ref] [n_raw] [mode] [jit]
Records are stored (16 B each) like in bench.py, except for the rare configurations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bridge_mc_b200 import bidding, compscore
from bridge_mc_b200._lib import TALLY_ALL, TALLY_SCORE
from bridge_mc_b200.engine import Constraints, Engine, seat_spec

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 26
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
jit = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = Engine(0)
mask = TALLY_ALL
p = bidding.openings.passes()
if cfg == "c1":
    cons = None
elif cfg == "c3":
    nt = bidding.nt_responses(1)
    cons = Constraints(None, [f"(and {p} {nt.chooses((3, 'NT'))})", p, bidding.openings.chooses((1, "NT")), p])
elif cfg == "c5":
    cons = Constraints(None, [bidding.nt_responses(2).chooses((6, "NT")), None, bidding.openings.chooses((2, "NT")), None])
elif cfg == "ref":
    cons = Constraints([seat_spec(hcp=(15, 17), s=(5, None)), seat_spec(hcp=(10, 11), h=(4, None))]).set_reference_order(2)
else:
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    if cfg == "c4":
        cs = [compscore.Contract(t).c_struct() for t in ("1NTS", "3NTS", "4HS", "4SS")] * 2
        cons.set_scoring(cs, [0, 0, 0, 0, 1, 1, 1, 1], [0, 1, 1, 2, 2, 3, 4, 5, 5, 6, 6, 7])
        mask |= TALLY_SCORE
if mode and cons is not None and cfg != "ref":
    cons.set_mode(mode)
if cons is not None and cfg != "ref":
    cons.set_jit(jit)
cap = n if cfg in ("c1", "ref") else n // 16
if cfg in ("c1", "ref"):
    cap = min(cap, 1 << 24)
for i in range(3):
    rec, tal, st = eng.sample(cons, n_raw=n, first_index=i << 40, cap=cap, tally_mask=mask, wave=n)      # one launch per call
    print(cfg, "raw", tal["raw"], "acc", tal["accepted"], "ms", round(st.kernel_ms, 3),
          "deals/s %.3e" % (tal["raw"] / st.kernel_ms * 1e3), "acc/s %.3e" % (tal["accepted"] / st.kernel_ms * 1e3),
          "mode", cons.mode if cons else 1)
