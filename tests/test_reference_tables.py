"""Outputs of the reference's own code (tests/golden/reference_tables.json, written by `oracle/run_reference.py tables`:
bridge.cl / bidding.cl / compscore.cl evaluated by oracle/minilisp.py) against the C oracle (no GPU needed) and, on the
GPU, against the product through the C ABI:

  contract-score  bridge.cl:3045-3069   the whole input domain, 3 360 values (the reference has no direct test of it)
  match-points    compscore.cl:123-131  the IMP scale for every |difference| 0..4100
  assess-hand / balanced? / choose-bid  bidding.cl:33-46, 146-150 on 1 500 seeded hands over `openings`,
                  `(nt-responses 1)`, `(nt-responses 2)` and `2C-responses` (the reference `eval`s the rule forms)
"""
import json
import os
import sys

import numpy as np
import pytest

from bridge_mc_b200 import sexp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import oracle as O  # noqa: E402

T = json.load(open(os.path.join(HERE, "golden", "reference_tables.json")))
SCHEMES = ["openings", "nt-responses-1", "nt-responses-2", "2c-responses"]


def scheme_bids(name):
    """the bids of a scheme's rows, in order: the car of every row of the reference's table"""
    rows = sexp.read(T["schemes"][name])
    return [[int(r[0][0]), str(r[0][1])] for r in rows]


def test_fixture_shape():
    cs = np.array(T["contract_score"])
    assert cs.shape == (5, 8, 3, 2, 14) and len(T["imps"]) == 4101 and len(T["hands"]) == 1500
    assert cs.min() == -7600 and cs.max() == 2980                     # the range of contract-score
    for name in SCHEMES:                                               # every row of every table fires on some hand
        seen = {json.dumps(h["bids"][name]) for h in T["hands"]}
        assert {json.dumps(b) for b in scheme_bids(name)} <= seen, name
    assert any(h["bids"]["openings"] is None for h in T["hands"])


def test_oracle_contract_score_is_the_reference_table():
    L = O.lib()
    cs = T["contract_score"]
    for suit in range(5):
        for level in range(8):
            for dbl in range(3):
                for vul in range(2):
                    for tricks in range(14):
                        assert L.orc_contract_score(suit, tricks, vul, dbl, level) == cs[suit][level][dbl][vul][tricks]


def test_oracle_imp_scale_is_the_reference_scale():
    for d, want in enumerate(T["imps"]):
        for sign in (1, -1):
            got = O.imps(sign * d)
            assert got == (None if want is None else sign * want if d else 0), d


def test_oracle_features_and_choose_bid_match_the_reference():
    bids = {name: scheme_bids(name) for name in SCHEMES}
    for h in T["hands"]:
        a = O.assess(h["hand"])
        assert a["lens"] == h["lens"] and a["power"] == h["power"] and a["losers"] == h["losers"] and a["balanced"] == h["balanced"]
        for sid, name in enumerate(SCHEMES):
            k = O.choose_bid(sid, h["hand"])
            want = h["bids"][name]
            assert (None if k < 0 else bids[name][k]) == want, (h["hand"], name)


# ---- the product (CUDA, through the C ABI) -------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_contract_score_is_the_reference_table(engine):
    from bridge_mc_b200._lib import Contract
    cs = np.array(T["contract_score"])
    for suit in range(5):
        for level in range(8):
            contracts = [Contract(level, suit, -1, dbl) for dbl in range(3) for _ in range(2)]
            vul = [v for _ in range(3) for v in range(2)]
            tricks = np.repeat(np.arange(14, dtype=np.uint8)[:, None], 6, axis=1)
            got = engine.score(contracts, vul, tricks)                  # [tricks, (dbl, vul)]
            assert np.array_equal(got.T.reshape(3, 2, 14), cs[suit, level])


@pytest.mark.gpu
def test_gpu_imp_scale_is_the_reference_scale(engine):
    d = np.arange(4101, dtype=np.int32)
    for sign in (1, -1):
        imps, status = engine.match_points(sign * d, np.zeros_like(d))
        for i, want in enumerate(T["imps"]):
            if want is None:
                assert status[i] == 1
            else:
                assert status[i] == 0 and imps[i] == sign * want


@pytest.mark.gpu
def test_gpu_features_and_choose_bid_match_the_reference(engine):
    from bridge_mc_b200 import bidding, cards
    from bridge_mc_b200.engine import Scheme
    tables = {"openings": bidding.openings, "nt-responses-1": bidding.nt_responses(1), "nt-responses-2": bidding.nt_responses(2),
              "2c-responses": bidding.two_c_responses}
    hands = T["hands"]
    # four of the fixture's hands per record would overlap: put each hand North of its own deal, the other 39 cards anywhere
    recs = []
    for h in hands:
        mine = {13 * s + r for s in range(4) for r in h["hand"][s]}
        rest = [c for c in range(52) if c not in mine]
        seats = [sorted(mine)] + [rest[13 * k:13 * k + 13] for k in range(3)]
        recs.append(cards.pack_deal([cards.Hand([[c % 13 for c in seat if c // 13 == s] for s in range(4)]) for seat in seats]))
    recs = np.array(recs, dtype=np.uint64)
    for name in SCHEMES:
        table = tables[name]
        bids = scheme_bids(name)
        _, feats = engine.filter(recs, scheme=Scheme(table.sexpr()))
        for i, h in enumerate(hands):
            f = feats[i]
            assert list(f["len"][0]) == h["lens"] and list(f["suit_hcp"][0]) == h["power"]
            assert f["losers"][0] == h["losers"] and bool(f["balanced"][0]) == h["balanced"]
            k = int(f["bid"][0])
            assert (None if k < 0 else bids[k]) == h["bids"][name], (h["hand"], name, k)
