"""A few small launches of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
The three shared-memory compaction queues of the seat-first sampler and the warp-cooperative list operations of the trick
estimator are the code a race detector is for."""
import sys

import numpy as np

sys.path.insert(0, ".")
from bridge_mc_b200 import bidding  # noqa: E402
from bridge_mc_b200.engine import Constraints, Engine, seat_spec  # noqa: E402

eng = Engine(0)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "sampler"):
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    rec, tal, _ = eng.sample(cons, n_raw=1 << 15, seed=3)
    print("mode-4 sampler", tal["accepted"])
    nt = bidding.nt_responses(1)
    p = bidding.openings.passes()
    c3 = Constraints(None, [f"(and {p} {nt.chooses((3, 'NT'))})", p, bidding.openings.chooses((1, "NT")), p])
    print("C3 interpreted", eng.sample(c3, n_raw=1 << 16, seed=3)[1]["accepted"])
    c1 = eng.sample(None, n_raw=1 << 12, seed=3)
    print("full deal", c1[1]["accepted"])
    ref = Constraints([seat_spec(hcp=(15, 17), s=(5, None)), seat_spec(hcp=(10, 11), h=(4, None))]).set_reference_order(2)
    print("reference order", eng.sample(ref, n_raw=1 << 10, seed=3)[1]["accepted"])
    _, feats = eng.filter(rec[:512], None, bidding.openings.scheme())
    print("filter", int(feats["hcp"].sum()))
if what in ("all", "playdeal"):
    rec = eng.sample(None, n_raw=16, seed=5)[0][:8]
    out = eng.play_deal(rec, -1, 0)
    print("play-deal", out["status"].tolist(), out["score"][:, 1].tolist())
