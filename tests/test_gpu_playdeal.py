"""The CUDA trick estimator (bmc_play_deal / bmc_play_hands, bridge_mc_b200/csrc/playdeal.cuh)
through the C ABI against (a) outputs of the reference's own `play-deal` (tests/golden/playdeal_ref.json, made by
oracle/run_reference.py running bridge.cl under oracle/minilisp.py), (b) the reference's inline play-deal test
(bridge.cl:2965, with the value its code computes), (c) the oracle restatement on seeded random deals, full deals
and endings.  Bit-exact: every card of every trick, the winners and the score."""
import json
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
REF = [r for r in json.load(open(os.path.join(HERE, "golden", "playdeal_ref.json")))["rows"] if "tricks" in r]


def lane_mask(suits):
    return sum(sum(1 << r for r in s) << (16 * i) for i, s in enumerate(suits))


def decode(row):
    n = int(row["n_tricks"])
    return ([[None if c == 255 else [int(c) >> 4, int(c) & 15] for c in row["cards"][t]] for t in range(n)],
            [int(w) for w in row["winner"][:n]], [int(row["score"][0]), int(row["score"][1])])


def rand_deal(rng, k):
    cards = rng.sample(range(52), 4 * k)
    return [[[c % 13 for c in sorted(cards[h * k:(h + 1) * k]) if c // 13 == s] for s in range(4)] for h in range(4)]


def test_matches_the_reference_run(engine):
    hands = np.array([[lane_mask(h) for h in r["hands"]] for r in REF], dtype=np.uint64)
    trump = np.array([-1 if r["trump"] is None else r["trump"] for r in REF], dtype=np.int8)
    out = engine.play_hands(hands, trump)
    assert (out["status"] == 0).all()
    for r, o in zip(REF, out):
        tricks, wno, score = decode(o)
        assert tricks == r["tricks"] and wno == r["winner_nos"] and score == r["score"]


def test_matches_the_oracle_on_random_deals(engine):
    from oracle import playdeal as P
    rng = random.Random(99)
    jobs = [(rand_deal(rng, k), rng.choice([-1, 0, 1, 2, 3])) for k in (1, 2, 3, 5, 7, 10) for _ in range(20)]
    jobs += [(rand_deal(rng, 13), rng.choice([-1, 0, 1, 2, 3])) for _ in range(24)]
    hands = np.array([[lane_mask(h) for h in hs] for hs, _ in jobs], dtype=np.uint64)
    out = engine.play_hands(hands, np.array([t for _, t in jobs], dtype=np.int8))
    for (hs, t), o in zip(jobs, out):
        ref = P.play_deal(None if t < 0 else t, *[P.Hand(h, id=i) for i, h in enumerate(hs)])
        assert o["status"] == 0
        tricks, wno, score = decode(o)
        assert tricks == ref.tricks and wno == ref.winner_nos and score == ref.score


def test_records_entry_point_rotates_to_the_declarer(engine):
    """bmc_play_deal on packed records = bmc_play_hands on the hands rotated so that the declarer comes first."""
    from bridge_mc_b200 import cards
    rng = random.Random(5)
    deals = [rand_deal(rng, 13) for _ in range(32)]
    rec = cards.pack_deals([[cards.Hand(h) for h in d] for d in deals])
    declarer = np.array([rng.randrange(4) for _ in deals], dtype=np.uint8)
    trump = np.array([rng.choice([-1, 0, 1, 2, 3]) for _ in deals], dtype=np.int8)
    a = engine.play_deal(rec, trump, declarer)
    hands = np.array([[lane_mask(d[(dec + k) % 4]) for k in range(4)] for d, dec in zip(deals, declarer)], dtype=np.uint64)
    b = engine.play_hands(hands, trump)
    assert (a["status"] == 0).all() and a.tobytes() == b.tobytes()
    bt = engine.best_tricks(rec, trump, declarer)
    assert (bt == a["score"][:, 1]).all()
    # scalar trump / declarer = the same for every deal
    c = engine.play_deal(rec, 2, 1)
    d = engine.play_deal(rec, np.full(len(deals), 2, np.int8), np.full(len(deals), 1, np.uint8))
    assert c.tobytes() == d.tobytes()


def test_host_mirror_play_deal_and_measures(engine):
    """The reference's inline full-deal test (bridge.cl:2965-2976) through the host mirror: the value the reference's
    code computes (tests/golden/reference_inline_tests.json; the literal in the file is stale), and msr-<contract>."""
    from bridge_mc_b200 import deal, playdeal, sexp
    from bridge_mc_b200.cards import str2hand
    inline = {r["line"]: r for r in json.load(open(os.path.join(HERE, "golden", "reference_inline_tests.json")))["results"]
              if r["file"] == "bridge.cl"}
    hands = [str2hand(t) for t in ("♣ AKQ7 ♦ 10 ♥ QJ632 ♠ 832", "♣ J54 ♦ KJ82 ♥ 8 ♠ KQJ95", "♣ 9832 ♦ A75 ♥ A975 ♠ A10",
                                  "♣ 106 ♦ Q9643 ♥ K104 ♠ 764")]
    out = playdeal.play_deal(None, *hands, engine=engine)
    want = sexp.read(inline[2965]["reference_computes"])
    assert out.tricks == [[list(c) for c in t] for t in want]
    assert playdeal.best_tricks(None, *hands, engine=engine) == out.score[1]
    assert playdeal.trump_length("H", *hands) == 9 and playdeal.longest_suit(*hands) == "H"
    with pytest.raises(playdeal.LispError):
        playdeal.best_tricks_star(None, *hands, engine=engine)          # the reference's own error (bridge.cl:3040)
    suit, in_suit, nt = playdeal.best_tricks_star(None, *hands, engine=engine, fixed=True)
    assert suit == "H" and nt == out.score[1]
    msr = deal.measure_parser(["3NTN"], engine=engine)[0]
    tricks, score = msr(None, *hands)
    assert tricks == out.score[1] and isinstance(score, int)


def test_arena_overflow_is_reported_not_silent(engine):
    rng = random.Random(3)
    hands = np.array([[lane_mask(h) for h in rand_deal(rng, 13)] for _ in range(8)], dtype=np.uint64)
    out = engine.play_hands(hands, -1, arena_kb=4)
    assert (out["status"] == 1).any() and (out["n_tricks"][out["status"] != 0] == 0).all()
    ok = engine.play_hands(hands, -1)
    assert (ok["status"] == 0).all()


def test_longest_expected_first_order_changes_nothing(engine, monkeypatch):
    """Large batches are played in the order of a probe pass's length estimate (play_deal_launch); the results must
    not depend on it: the same 1 200 deals with the order forced on, forced off, and with small arenas + retry."""
    rng = random.Random(17)
    hands = np.array([[lane_mask(h) for h in rand_deal(rng, 13)] for _ in range(1200)], dtype=np.uint64)
    trump = np.array([rng.choice([-1, 0, 1, 2, 3]) for _ in range(1200)], dtype=np.int8)
    monkeypatch.setenv("BMC_PD_ORDER", "0")
    plain = engine.play_hands(hands, trump)
    monkeypatch.setenv("BMC_PD_ORDER", "1")
    ordered = engine.play_hands(hands, trump)
    monkeypatch.setenv("BMC_PD_ARENA_KB", "64")                   # many deals outgrow 64 KB: the retry pass runs too
    retried = engine.play_hands(hands, trump)
    assert (plain["status"] == 0).all()
    assert plain.tobytes() == ordered.tobytes() == retried.tobytes()


def test_outcomes_of_the_frozen_deal_set(engine):
    """1 500 seeded deals and endings with random trumps against the digest frozen in tests/test_playdeal_oracle.py."""
    from test_playdeal_oracle import FROZEN_DIGEST, frozen_deals, frozen_digest
    jobs = frozen_deals()
    hands = np.array([[sum(int(m[4 * j + s]) << (16 * s) for s in range(4)) for j in range(4)] for _, m in jobs], dtype=np.uint64)
    out = engine.play_hands(hands, np.array([t for t, _ in jobs], dtype=np.int8))
    assert (out["status"] == 0).all()
    rows = [(0, int(o["score"][0]), int(o["score"][1]), int(o["n_tricks"]), [[int(c) for c in o["cards"][t]] for t in range(int(o["n_tricks"]))],
             [int(w) for w in o["winner"][:int(o["n_tricks"])]]) for o in out]
    assert frozen_digest(rows) == FROZEN_DIGEST


def test_edge_cases_follow_the_reference(engine):
    """hands of unequal size: the reference signals a Lisp error when an empty hand has to play (status 2); a longer
    declarer hand just keeps its extra card; four empty hands: play-deal returns nil (status 4) -- as the oracle does"""
    from oracle import playdeal as P
    cases = [[[[0], [], [], []], [[], [], [], []], [[1], [], [], []], [[2], [], [], []]],
             [[[0, 1], [], [], []], [[2], [], [], []], [[3], [], [], []], [[4], [], [], []]],
             [[[], [], [], []]] * 4]
    out = engine.play_hands(np.array([[lane_mask(h) for h in hs] for hs in cases], dtype=np.uint64), -1)
    assert out["status"].tolist() == [2, 0, 4]
    with pytest.raises(P.LispError):
        P.play_deal(None, *[P.Hand(h, id=i) for i, h in enumerate(cases[0])])
    ref = P.play_deal(None, *[P.Hand(h, id=i) for i, h in enumerate(cases[1])])
    tricks, wno, score = decode(out[1])
    assert tricks == ref.tricks and score == ref.score
    assert P.play_deal(None, *[P.Hand(h, id=i) for i, h in enumerate(cases[2])]) is None
    assert engine.play_hands(np.empty((0, 4), dtype=np.uint64), -1).size == 0
