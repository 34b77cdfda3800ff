"""Driver for profiling K2/K4: bmc_filter with a scheme, bmc_score, bmc_match_points, bmc_score_tally."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bridge_mc_b200 import bidding, compscore
from bridge_mc_b200.engine import Constraints, Engine
eng = Engine(0)
rec, _, _ = eng.sample(None, n_raw=1 << 22, seed=1)
cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
for _ in range(2):
    acc, feats = eng.filter(rec, cons, bidding.openings.scheme())
for _ in range(2):
    eng.filter(rec, None, None)                      # features only: the HBM-bound form of K2
cs = [compscore.Contract(t).c_struct() for t in ("1NTS", "3NTS", "4HS", "4SS")]
tr = np.random.default_rng(1).integers(0, 14, size=(1 << 22, 4)).astype(np.uint8)
for _ in range(2):
    sc = eng.score(cs, [0, 1, 0, 1], tr)
    imps, st = eng.match_points(sc[:, 0], sc[:, 1])
    eng.score_tally(cs * 2, [0, 0, 0, 0, 1, 1, 1, 1], [0, 1, 2, 3], n=1 << 26)
print("records", rec.shape[0], "accepted", int(acc.sum()))
