"""Sampling DISTRIBUTION of the reference's own `build` (bridge.cl:1292-1319), drawn by running the reference's source
under oracle/minilisp.py (tests/golden/reference_build_samples.json, `oracle/run_reference.py build`): two fixed hands and
two ranged seats -- all four seats defined, so `infer-distgen-params` infers lower bounds too, and the last ranged seat is
never sampled (bridge.cl:1303).  The reference holds no test of a sampling distribution; these
samples pin (a) the oracle's restatement of distgen + infer-distgen-params + build (oracle/ref_build.py, no GPU needed)
and (b) the reference-order sampler on the GPU (mode 3), by two-sample chi-square on every seat's HCP and suit lengths."""
import json
import os
import sys

import numpy as np
import pytest
from scipy.stats import chi2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import oracle as O  # noqa: E402
from oracle.ref_build import ref_build_batch  # noqa: E402

G = json.load(open(os.path.join(HERE, "golden", "reference_build_samples.json")))


def _deals():
    fixed = G["fixed"]
    taken = sum(1 << (13 * s + r) for h in fixed for s in range(4) for r in h[s])
    out = []
    for m in G["third_hand"]:
        rest = ((1 << 52) - 1) & ~taken & ~m
        out.append(fixed + [[[r for r in range(13) if (x >> (13 * s + r)) & 1] for s in range(4)] for x in (m, rest)])
    return out


DEALS = _deals()
FIXED = [[[9, 10, 11, 12], [9, 10, 11, 12], [0, 1, 2], [0, 1]],       # ((A K Q J) (A K Q J) (4 3 2) (3 2)) as ranks
         [[6, 7, 8], [6, 7, 8], [3, 11, 12], [2, 3, 11, 12]]]         # ((10 9 8) (10 9 8) (A K 5) (A K 5 4))
DEFS = [dict(hcp=(2, 4), h=(3, None)), dict(hcp=(2, 4), s=(2, None))]
POWER = [0] * 9 + [1, 2, 3, 4]


def feats(hand):
    return sum(POWER[r] for s in hand for r in s), [len(s) for s in hand]


def ref_hists():
    """per seat (build's result order): HCP histogram [38] and suit-length histograms [4][14]"""
    hcp = np.zeros((4, 38), int)
    ln = np.zeros((4, 4, 14), int)
    for d in DEALS:
        for k, h in enumerate(d):
            p, lens = feats(h)
            hcp[k, p] += 1
            for s in range(4):
                ln[k, s, lens[s]] += 1
    return hcp, ln


def two_sample_p(obs, pmf):
    """chi-square of the observed counts against a pmf estimated from a much larger sample; bins merged to expected >= 5"""
    obs, exp = np.asarray(obs, float), np.asarray(pmf, float) * np.sum(obs)
    o, e, mo, me = [], [], 0.0, 0.0
    for i in np.argsort(exp):
        mo += obs[i]; me += exp[i]
        if me >= 5:
            o.append(mo); e.append(me); mo = me = 0.0
    if me > 0 and e:
        o[-1] += mo; e[-1] += me
    if len(o) < 2:
        return 1.0
    o, e = np.array(o), np.array(e)
    return chi2.sf(((o - e) ** 2 / e).sum(), len(o) - 1)


def test_reference_samples_satisfy_their_own_defs():
    assert len(DEALS) >= 20000
    for d in DEALS:
        assert [[sorted(s) for s in h] for h in d[:2]] == FIXED and len(d) == 4
        cards = sorted(13 * s + r for h in d for s in range(4) for r in h[s])
        assert cards == list(range(52))
        p2, l2 = feats(d[2])
        p3, l3 = feats(d[3])
        assert 2 <= p2 <= 4 and l2[2] >= 3
        assert 2 <= p3 <= 4          # inferred: the six points left over must split 2..4 / 2..4
    # the never-sampled last seat: its spade range reaches it only through the narrowing of the third seat -- whose open
    # upper ends make the inferred LOWER spade bound the whole supply (bridge.cl:1101-1104), so the split is always 5 - 2
    assert all(len(d[2][3]) == 5 and len(d[3][3]) == 2 for d in DEALS)


def test_oracle_build_matches_the_reference_samples():
    fixed52 = [sum(1 << (13 * s + r) for s in range(4) for r in h[s]) for h in FIXED]
    hands, n_err = ref_build_batch(fixed52, DEFS, 400_000, 3)
    assert n_err == 0
    lanes = O.mask52_to_lanes64(hands.reshape(-1)).reshape(-1, 4)
    f = O.features_batch(lanes.reshape(-1)).reshape(lanes.shape[0], 4, 11)
    hcp, ln = ref_hists()
    ps = []
    for k in (2, 3):
        ps.append(two_sample_p(hcp[k], np.bincount(f[:, k, 8], minlength=38)[:38] / f.shape[0]))
        for s in range(4):
            ps.append(two_sample_p(ln[k, s], np.bincount(f[:, k, s], minlength=14)[:14] / f.shape[0]))
    assert min(ps) > 1e-4, ps                      # 10 tests


@pytest.mark.gpu
def test_gpu_reference_order_matches_the_reference_samples(engine):
    from bridge_mc_b200.cards import Hand
    from bridge_mc_b200.engine import Constraints, seat_spec
    cons = Constraints([seat_spec(fixed=Hand(h)) for h in FIXED] + [seat_spec(**d) for d in DEFS]).set_reference_order(2)
    n = 4_000_000
    _, tal, _ = engine.sample(cons, n_raw=n, seed=21, cap=0)
    assert tal["accepted"] + tal["voided"] == n and tal["accepted"] == n
    hcp, ln = ref_hists()
    ps = []
    for k in (2, 3):
        ps.append(two_sample_p(hcp[k], tal["hcp"][k] / tal["accepted"]))
        for s in range(4):
            ps.append(two_sample_p(ln[k, s], tal["len"][k][s] / tal["accepted"]))
    assert min(ps) > 1e-4, ps
