"""Large-sample exactness check of the samplers (not part of the test suite): chi-square of the device tallies
against exact combinatorial pmfs at 1e11 raw deals.

  full-deal kernel  : unconstrained, every seat's HCP (38 bins) and sorted shape (39 bins)
  seat-first kernel : South test-bid (1 NT); acceptance against the exact probability (times the exact
                      survival factors of the void checks) and South's 15/16/17 split; then, with an all-admitting
                      window, every seat's HCP / shape and South's suit lengths through the rank / unrank path
"""
import json, os, sys
from math import comb, prod
from itertools import permutations
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.stats import chi2, norm
from bridge_mc_b200 import bidding
from bridge_mc_b200._lib import shape_table
from bridge_mc_b200.engine import Constraints, Engine

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e11)
eng = Engine(0)


def chi2_p(obs, exp):
    obs, exp = np.asarray(obs, float), np.asarray(exp, float)
    keep = exp >= 5
    o = np.append(obs[keep], obs[~keep].sum())
    e = np.append(exp[keep], exp[~keep].sum())
    if e[-1] < 5:
        o[-2] += o[-1]; e[-2] += e[-1]; o, e = o[:-1], e[:-1]
    stat = ((o - e) ** 2 / e).sum()
    return float(stat), len(o) - 1, float(chi2.sf(stat, len(o) - 1))


# exact unconstrained pmfs
hcp_num = np.zeros(38)
for a in range(5):
    for k in range(5):
        for q in range(5):
            for j in range(5):
                if a + k + q + j <= 13:
                    hcp_num[4 * a + 3 * k + 2 * q + j] += comb(4, a) * comb(4, k) * comb(4, q) * comb(4, j) * comb(36, 13 - a - k - q - j)
total = comb(52, 13)
assert hcp_num.sum() == total
shape_num = np.array([sum(comb(13, p[0]) * comb(13, p[1]) * comb(13, p[2]) * comb(13, p[3]) for p in set(permutations(r)))
                      for r in shape_table().tolist()], float)
out = {"raw": N}
_, t, st = eng.sample(None, n_raw=N, seed=20261020, cap=0)
acc = t["accepted"]
out["full_deal"] = {"accepted": acc, "voided": t["voided"], "kernel_ms": st.kernel_ms,
                    "hcp": [chi2_p(t["hcp"][k], hcp_num / total * acc) for k in range(4)],
                    "shape": [chi2_p(t["shape"][k], shape_num / total * acc) for k in range(4)]}

cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None]).set_mode(2)
_, t, st = eng.sample(cons, n_raw=N, seed=20261021, cap=0)
surv = (2 ** 64 // total) * total / 2 ** 64            # stage A1: u beyond S * C(52,13); stage A2 draws no random numbers
surv *= prod(1 - ((1 << 32) % prod(m for m in range(39 - 4 * w - 3, 39 - 4 * w + 1) if m >= 2)) / 2 ** 32 for w in range(10))
p = 24235203726 / total * surv
z = (t["accepted"] - N * p) / (N * p * (1 - p)) ** 0.5
split = np.array([9820068156, 8126568540, 6288567030], float)
out["seat_first_1nt"] = {"accepted": t["accepted"], "expected": N * p, "z": z, "p_two_sided": float(2 * norm.sf(abs(z))),
                         "kernel_ms": st.kernel_ms,
                         "south_hcp_split": chi2_p(t["hcp"][2][15:18], split / split.sum() * t["accepted"])}
# the whole rank / unrank path at acceptance 1 (an all-admitting window on South), N / 10 raw deals
from bridge_mc_b200.engine import seat_spec
cons = Constraints([None, None, seat_spec(hcp=(0, 37)), None]).set_mode(2)
_, t, st = eng.sample(cons, n_raw=N // 10, seed=20261022, cap=0)
acc = t["accepted"]
len_num = np.array([comb(13, l) * comb(39, 13 - l) for l in range(14)], float)
out["seat_first_all_hands"] = {"raw": N // 10, "accepted": acc, "voided": t["voided"], "kernel_ms": st.kernel_ms,
                               "hcp": [chi2_p(t["hcp"][k], hcp_num / total * acc) for k in range(4)],
                               "shape": [chi2_p(t["shape"][k], shape_num / total * acc) for k in range(4)],
                               "south_len": [chi2_p(t["len"][2][su], len_num / total * acc) for su in range(4)]}
print(json.dumps(out, indent=1))
