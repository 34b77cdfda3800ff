"""GPU tests of the reference-facing host API (the Python mirror of the Lisp entry
points) and of the remaining C-ABI calls; checked against the reference's own
KATs and the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from kat_util import bid_of, hand_of, load_kats, truth, unquote
from bridge_mc_b200 import bidding, cards, compscore, hist
from bridge_mc_b200._lib import BadRange, Contract as CContract
from bridge_mc_b200.cards import seat_masks, str2hand
from bridge_mc_b200.deal import Deal, McCase, mc_deal
from bridge_mc_b200.engine import Constraints, seat_spec
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def sort_records(rec):
    rec = np.ascontiguousarray(rec, dtype=np.uint64).reshape(-1, 2)
    return rec[np.lexsort((rec[:, 0], rec[:, 1]))]


# ---- bidding.cl KATs through the GPU path --------------------------------------------
def _scheme(form):
    if len(form) == 2:
        return bidding.openings
    s = form[2]
    if s[0] == "NT-RESPONSES":
        return bidding.nt_responses(s[1])
    if s == "2C-RESPONSES":
        return bidding.two_c_responses
    if s[0] == "BASIC-RESPONSES":
        return bidding.basic_responses(bid_of(s[1]))
    if s[0] == "FURTHER-BID":
        return bidding.further_bid(unquote(s[1]), unquote(s[2]), bid_of(s[3]))
    return None


def test_bidding_kats_on_gpu(engine):
    n = 0
    for where, form, exp in load_kats("SUIT-LOOSERS"):
        pass
    for where, form, exp in load_kats("ASSESS-HAND"):
        assert bidding.balanced(hand_of(form[1]), engine) == truth(exp), where
        n += 1
    for where, form, exp in load_kats("TEST-BID"):
        assert bidding.test_bid(bid_of(form[1]), hand_of(form[2]), engine=engine) == truth(exp), where
        n += 1
    for where, form, exp in load_kats("CHOOSE-BID"):
        table = _scheme(form)
        assert table is not None, where
        assert bidding.choose_bid(hand_of(form[1]), table, engine) == bid_of(exp), where
        n += 1
    assert n == 6 + 28 + 10 + 13 + 8 + 7 + 6


def test_bidding_class_next(engine):
    """bidding.cl:482-487: the reference's own example deal; North (seat 0) opens 1D"""
    deal = [str2hand("N: ♠ Q1097 ♥ 3 ♦ KJ1063 ♣ K102"), str2hand("E: ♠ 62 ♥ AJ75 ♦ A942 ♣ 987"),
            str2hand("S: ♠ AKJ5 ♥ K1064 ♦ Q87 ♣ A6"), str2hand("W: ♠ 843 ♥ Q982 ♦ 5 ♣ QJ543")]
    b = bidding.Bidding(deal, engine)
    bids = b.next()
    ref = [O.choose_bid(0, h) for h in deal]
    first = next(i for i, k in enumerate(ref) if k >= 0)
    assert bids == [bidding.openings.rows[ref[first]][0]]
    assert b.deal[0] == deal[first]
    assert b.bid_scheme is not None and b.bid_scheme.bids()


def test_assess_hand_matches_oracle(engine):
    h = str2hand("♣ KJ9753 ♦ 94 ♥ A84 ♠ A9")
    lens, power, losers = bidding.assess_hand(h, engine=engine)
    ref = O.assess(h)
    assert (lens, power, losers) == (ref["lens"], ref["power"], ref["losers"])
    assert bidding.assess_hand(h, lambda l, p, lo: sum(p), engine) == ref["hcp"]


def test_unspliced_rule_quirk(engine):
    """bidding.cl:249: the (3 suit) raise holds an un-spliced list; reaching it is an error in the
    reference, and the club raise is simply NIL."""
    t = bidding.basic_responses((1, "H"))
    h = str2hand("♣ 72 ♦ K854 ♥ K653 ♠ AQ8")        # 12 hcp, 4 hearts, no 4 spades: reaches the (3 H) row
    with pytest.raises(bidding.BidError):
        bidding.choose_bid(h, t, engine)
    tc = bidding.basic_responses((1, "C"))
    h2 = str2hand("♣ KQ8542 ♦ K85 ♥ K6 ♠ 98")       # 11 hcp, club fit, no 4-card suit outside
    assert bidding.choose_bid(h2, tc, engine) is None


# ---- compscore.cl KATs through the product API ------------------------------------------
def test_compscore_kats_on_gpu(engine):
    for where, form, exp in load_kats("BASE-SCORE"):
        c = compscore.Contract(str(form[1][3]))
        vul = len(form) > 2 and truth(form[3])
        assert compscore.base_score(c, vulnerable=vul, engine=engine) == exp, where
    for where, form, exp in load_kats("MATCH-POINTS"):
        a, b = compscore.Contract(str(form[1][3])), compscore.Contract(str(form[2][3]))
        vul = str(unquote(form[4])).lower() if len(form) > 3 else None
        assert compscore.match_points(a, b, vulnerable=vul, engine=engine) == exp, where
    assert compscore.contract_score("H", 10, engine=engine) == 420
    assert compscore.contract_score(None, 9, vulnerable=True, level=3, engine=engine) == 600
    with pytest.raises(ValueError):
        compscore.match_points(compscore.Contract("7NTNxx="), compscore.Contract("7NTNxx-13"), vulnerable="both",
                               engine=engine)


# ---- reference-generated deals -> GPU filter ---------------------------------------------
def test_filter_on_reference_generated_deals(engine):
    """Deals produced by the reference's own generator (distgen exact sampler restated in the
    oracle) go through bmc_filter: accept flags and features must be bit-exact."""
    g = O.Distgen(hcp=(15, 17), c=(2, 5), d=(2, 5), h=(2, 4), s=(2, 4))
    hands = g.build_batch(3000, seed=42)                  # seat 0 = the ranged seat
    lanes = O.mask52_to_lanes64(hands)
    rec = np.zeros((hands.shape[0], 2), dtype=np.uint64)
    for k in range(4):
        if k & 1:
            rec[:, 0] |= lanes[:, k]
        if k & 2:
            rec[:, 1] |= lanes[:, k]
    cons = Constraints([seat_spec(hcp=(15, 17), c=(2, 5), d=(2, 5), h=(2, 4), s=(2, 4)), None, None, None],
                       [None, bidding.openings.passes(), None, None])
    acc, feats = engine.filter(rec, cons, bidding.openings.scheme())
    f = O.features_batch(lanes.reshape(-1)).reshape(-1, 4, 11)
    bids = O.choose_bid_batch(0, lanes.reshape(-1)).reshape(-1, 4)
    assert np.array_equal(feats["hcp"].astype(np.int32), f[:, :, 8])
    assert np.array_equal(feats["len"].astype(np.int32), f[:, :, 0:4])
    assert np.array_equal(feats["losers"].astype(np.int32), f[:, :, 9])
    assert np.array_equal(feats["bid"].astype(np.int32), bids)
    assert np.array_equal(acc.astype(bool), bids[:, 1] == -1)      # seat 0 always inside its ranges


def test_multi_seat_constraint_accept_flags(engine):
    """C3-style: N, E, W pass, S opens 1NT, N answers 3NT to 1NT (bidding.cl:76-106,166-187)."""
    nt = bidding.nt_responses(1)
    p = bidding.openings.passes()
    cons = Constraints(None, [f"(and {p} {nt.chooses((3, 'NT'))})", p, bidding.openings.chooses((1, "NT")), p])
    rec, _, _ = engine.sample(None, n_raw=400_000, seed=31)
    acc, _ = engine.filter(rec, cons, None, want_features=False)
    m = seat_masks(rec)
    op = O.choose_bid_batch(0, m.reshape(-1)).reshape(-1, 4)
    resp = O.choose_bid_batch(1, m[:, 0])
    ref = (op[:, 0] == -1) & (op[:, 1] == -1) & (op[:, 2] == 2) & (op[:, 3] == -1) & (resp == 6)
    assert np.array_equal(acc.astype(bool), ref)
    assert 200 < ref.sum() < 900                         # p ~ 1.26e-3
    cons.set_mode(1)
    got, tal, _ = engine.sample(cons, n_raw=400_000, seed=31)
    assert np.array_equal(sort_records(got), sort_records(rec[ref]))


# ---- fixed (:=) hands ------------------------------------------------------------------------
def test_fixed_hand_sampler_bit_exact(engine):
    h = str2hand("♣ AJ7 ♦ AK54 ♥ 6543 ♠ 64")
    h2 = str2hand("♣ KQ2 ♦ QJ ♥ AKQ2 ♠ AK32")
    for fixed in ([0, h.mask64(), 0, 0], [h2.mask64(), 0, 0, h.mask64()], [h.mask64(), h2.mask64(), 0, 0]):
        specs = [seat_spec(fixed=m) if m else None for m in fixed]
        cons = Constraints(specs)
        n = 50_000
        rec, tal, _ = engine.sample(cons, n_raw=n, seed=77, first_index=9)
        ref, void = O.gpu_deal_fixed_batch(77, 9, n, fixed)
        assert tal["voided"] == int(void.sum())
        assert np.array_equal(sort_records(rec), sort_records(ref[void == 0]))
        m = seat_masks(rec)
        for k, fm in enumerate(fixed):
            if fm:
                assert (m[:, k] == np.uint64(fm)).all()


def test_build_property_kat(engine):
    """bridge.cl:2989-3003: three ranged seats + one fixed hand; every generated hand (incl. the
    never-directly-sampled last seat) has HCP <= 11 and all suit lengths in 2..5."""
    rng = {"hcp": [0, 11], "c": [2, 5], "d": [2, 5], "h": [2, 5], "s": [2, 5]}
    fixed = cards.Hand([[12, 9, 5], [12, 11, 3, 2], [4, 3, 2, 1], [4, 2]])     # ((A J 7) (A K 5 4) (6 5 4 3) (6 4))
    dg = Deal([fixed], [rng, rng, rng], engine=engine, seed=5)
    seen = set()
    for _ in range(20):
        hands = dg.build()
        assert len(hands) == 4 and hands[0] == fixed
        seen.add(tuple(tuple(s) for s in hands[1].suits))
        for h in hands[1:]:
            hcp = sum(max(r - 8, 0) for s in h.suits for r in s)
            assert hcp <= 11 and all(2 <= len(s) <= 5 for s in h.suits)
        assert sorted(c for h in hands for c in h.cards()) == list(range(52))
    assert len(seen) == 20


def test_deal_build_order_and_errors(engine):
    d = Deal(defs=["(:hcp (15 17) :s (5 nil))"], engine=engine, seed=1)
    hands = d.build()
    assert len(hands) == 4
    assert 15 <= sum(max(r - 8, 0) for s in hands[0].suits for r in s) <= 17 and len(hands[0].suits[3]) >= 5
    three = [str2hand("♣ AKQJ1098765432 ♦ ♥ ♠"), str2hand("♣ ♦ AKQJ1098765432 ♥ ♠"), str2hand("♣ ♦ ♥ AKQJ1098765432 ♠")]
    wide = {"hcp": [10, 10], "c": [0, 13], "d": [0, 13], "h": [0, 13], "s": [0, 13]}
    only = Deal(three, [wide], engine=engine, seed=2).build()
    assert len(only) == 1 and only[0].suits[3] == list(range(13))      # bridge.cl:1316-1319: only the sampled hand
    # kept quirk: a suit without an explicit upper bound is capped at its low cards (bridge.cl:1000-1001), so the
    # reference cannot build this 13-spade remainder at all (total weight 0 -> error); uniform semantics can
    with pytest.raises(BadRange):
        Deal(three, [{"hcp": [10, 10]}], engine=engine, seed=2).build()
    assert len(Deal(three, [{"hcp": [10, 10]}], engine=engine, seed=2, semantics="uniform").build()) == 1
    with pytest.raises(BadRange):
        Deal(defs=[{"hcp": [30, 37]}, {"hcp": [15, 17]}], engine=engine).build()


def test_mc_case_and_mc_deal(engine):
    def south_hcp(trump, n, e, s, w):
        return sum(max(r - 8, 0) for suit in s.suits for r in suit)

    def ns_spades(trump, n, e, s, w):
        return len(n.suits[3]) + len(s.suits[3])

    deal = Deal(defs=[{"hcp": [15, 17], "c": [2, 5], "d": [2, 5], "h": [2, 4], "s": [2, 4]}], engine=engine, seed=3)
    mc = McCase(deal, [south_hcp, ns_spades], reorder=[1, 2, 0, 3])      # the ranged seat becomes South
    rows = mc.sim(500)
    assert len(rows) == 500 and all(15 <= r[1] <= 17 for r in rows)
    h = hist.histogram([r[1] for r in rows])
    assert sum(e[1] for e in h) == 1 and {e[0] for e in h} == {15, 16, 17}
    assert len(mc.select(fields=["south_hcp"], filter=lambda r: r["south_hcp"] == 15)) == sum(r[1] == 15 for r in rows)
    mc.add_prop(lambda trump, n, e, s, w: len(s.suits[0]))
    assert all(2 <= r[3] <= 5 for r in mc.data)
    # rest.cl:264-277: fixed South hand + ranged North, request order N E S W
    south = str2hand("S: ♠ AKJ5 ♥ K1064 ♦ Q87 ♣ A6")
    mc2, tree = mc_deal([{"hcp": [10, 12], "s": [4, None]}, {}, south, {}], [ns_spades], runs=200, engine=engine, seed=4)
    assert all(r[0][2] == south for r in mc2.data)
    assert all(len(r[0][0].suits[3]) >= 4 for r in mc2.data)
    assert sum(e[1] for e in tree) == 1


def test_mc_case_save_load(engine, tmp_path):
    deal = Deal(engine=engine, seed=8)
    mc = McCase(deal, [lambda t, n, e, s, w: len(n.suits[0])])
    mc.sim(50)
    path = str(tmp_path / "case.dat")
    mc.save(path)
    back = McCase.load(path)
    assert [[h.suits for h in r[0]] for r in back.data] == [[h.suits for h in r[0]] for r in mc.data]
    assert [r[1] for r in back.data] == [r[1] for r in mc.data]


# ---- compscore tallies ----------------------------------------------------------------------------------
SCORE_CONTRACTS = ("1NTS", "3NTS", "4HS", "4SS", "1NTS", "3NTS", "4HS", "7NTNxx")
SCORE_VUL = [0, 0, 0, 0, 1, 1, 1, 1]
SCORE_PAIRS = [0, 1, 1, 2, 2, 3, 4, 5, 5, 6, 6, 7, 7, 3]           # the last two pairs overflow the IMP table at times


def test_score_tally_bit_exact_vs_oracle(engine):
    """bmc_score_tally (index-keyed trick source -> base-score -> match-points) against the CPU mirror, word for word:
    trick histogram, IMP histogram and the count of differences off the IMP table (the reference errors there)"""
    cs = [compscore.Contract(t).c_struct() for t in SCORE_CONTRACTS]
    n = 300_000
    th, ih, tab = engine.score_tally(cs, SCORE_VUL, SCORE_PAIRS, n=n, seed=9, first_index=12345678901)
    tr_o, im_o, dr_o = O.score_tally_index(9, 12345678901, n, cs, SCORE_VUL, SCORE_PAIRS)
    assert np.array_equal(th.astype(np.int64), tr_o[:len(cs)])
    assert np.array_equal(ih.astype(np.int64), im_o[:len(SCORE_PAIRS) // 2])
    assert np.array_equal(engine.last_imp_dropped.astype(np.int64), dr_o[:len(SCORE_PAIRS) // 2])
    assert dr_o[5:7].sum() > 0                                            # the overflow path was exercised
    assert (th.sum(axis=1) == n).all() and (ih.sum(axis=1) + engine.last_imp_dropped == n).all()
    exp = n / 14
    assert (abs(th.astype(float) - exp) < 6 * exp ** 0.5).all()          # exactly uniform tricks 0..13
    for c in range(len(cs)):
        for t in range(14):
            assert tab[c, t] == O.lib().orc_base_score(cs[c].level, cs[c].suit, cs[c].declarer, cs[c].dbl, t, SCORE_VUL[c])
    # additivity over index sub-ranges (what lets GPUs split a job)
    a = engine.score_tally(cs, SCORE_VUL, SCORE_PAIRS, n=200_000, seed=9, first_index=12345678901)
    b = engine.score_tally(cs, SCORE_VUL, SCORE_PAIRS, n=100_000, seed=9, first_index=12345678901 + 200_000)
    assert np.array_equal(a[0] + b[0], th) and np.array_equal(a[1] + b[1], ih)


@pytest.mark.parametrize("mode", [1, 2])
def test_fused_score_tally_over_constrained_sample(engine, mode):
    """BASELINE config 4: compscore tallies over CONSTRAINED samples.  The sampling kernel scores every accepted deal
    in stage B (BMC_TALLY_SCORE); the result must equal the CPU mirror applied to the records the same call stored."""
    from bridge_mc_b200._lib import TALLY_ALL, TALLY_SCORE
    cs = [compscore.Contract(t).c_struct() for t in SCORE_CONTRACTS]
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None]).set_mode(mode)
    cons.set_scoring(cs, SCORE_VUL, SCORE_PAIRS)
    n = 1 << 22
    rec, tal, _ = engine.sample(cons, n_raw=n, seed=21, tally_mask=TALLY_ALL | TALLY_SCORE)
    assert rec.shape[0] == tal["accepted"] > 100_000
    tr_o, im_o, dr_o = O.score_tally_records(21, rec, cs, SCORE_VUL, SCORE_PAIRS)
    assert np.array_equal(tal["tricks"], tr_o) and np.array_equal(tal["imps"], im_o) and np.array_equal(tal["imp_dropped"], dr_o)
    assert (tal["tricks"].sum(axis=1) == tal["accepted"]).all()
    # the tally is a function of the accepted SET: any partition of the index range gives the same words
    parts = [engine.sample(cons, n_raw=n // 4, seed=21, first_index=i * (n // 4), cap=0, tally_mask=TALLY_SCORE)[1] for i in range(4)]
    for key in ("tricks", "imps", "imp_dropped"):
        assert np.array_equal(tal[key], sum(p[key] for p in parts))
    # without the bit nothing is scored; without contracts the bit is an error
    _, plain, _ = engine.sample(cons, n_raw=1 << 16, seed=21, cap=0)
    assert plain["tricks"].sum() == 0 and plain["imps"].sum() == 0
    from bridge_mc_b200._lib import BridgeMcError
    with pytest.raises(BridgeMcError):
        engine.sample(Constraints(None, [None, None, bidding.openings.form((1, "NT")), None]), n_raw=1 << 16, cap=0,
                      tally_mask=TALLY_SCORE)


# ---- C5: the rare constraint (S chooses 2NT, N answers 6NT over (nt-responses 2)) ---------------------------
def _c5_rules():
    return [bidding.nt_responses(2).chooses((6, "NT")), None, bidding.openings.chooses((2, "NT")), None]


def _c5_oracle(rec):
    m = seat_masks(rec)
    return (O.choose_bid_batch(0, m[:, 2]) == 1) & (O.choose_bid_batch(2, m[:, 0]) == 6)     # openings[1] = 2NT, nt2[6] = 6NT


def test_c5_constraint_parity(engine):
    """bidding.cl:81-82,166-187.  (i) both dealing modes: every accepted deal satisfies the oracle's choose-bid, and the
    acceptance agrees with the known 2.1685e-5; (ii) on deals where South does open 2NT the accept flag is bit-exact
    against the oracle (near misses included); (iii) the full-deal sampler's accepted set = the filter of the
    unconstrained stream over the same indices."""
    n = 100_000_000
    counts = []
    for mode in (1, 2):
        cons = Constraints(None, _c5_rules()).set_mode(mode)
        rec, tal, _ = engine.sample(cons, n_raw=n, seed=55, cap=8192)
        assert rec.shape[0] == tal["accepted"] and _c5_oracle(rec).all(), mode
        counts.append(tal["accepted"])
        exp = 2.1685e-5 * n
        assert abs(tal["accepted"] - exp) < 6 * exp ** 0.5, (mode, tal["accepted"])
    south_2nt = Constraints(None, [None, None, bidding.openings.chooses((2, "NT")), None])
    near, _, _ = engine.sample(south_2nt, n_raw=12_000_000, seed=56)
    assert near.shape[0] > 50_000
    acc, _ = engine.filter(near, Constraints(None, _c5_rules()), None, want_features=False)
    ref = _c5_oracle(near)
    assert np.array_equal(acc.astype(bool), ref) and 100 < ref.sum() < 600
    raw, _, _ = engine.sample(None, n_raw=4_000_000, seed=57)
    got, _, _ = engine.sample(Constraints(None, _c5_rules()).set_mode(1), n_raw=4_000_000, seed=57)
    assert np.array_equal(sort_records(got), sort_records(raw[_c5_oracle(raw)]))


def test_sampler_partition_invariance(engine):
    """tallies of [0,N) equal the sum over any partition into index ranges (what makes the multi-GPU
    all-reduce bit-identical for every G)"""
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    n = 1 << 20
    _, whole, _ = engine.sample(cons, n_raw=n, seed=3, cap=0)
    parts = [engine.sample(cons, n_raw=n // 4, seed=3, first_index=i * (n // 4), cap=0)[1] for i in range(4)]
    for key in ("raw", "voided", "accepted"):
        assert whole[key] == sum(p[key] for p in parts)
    for key in ("hcp", "len", "shape"):
        assert np.array_equal(whole[key], sum(p[key] for p in parts))


def test_accept_target_mode(engine):
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    rec, tal, st = engine.sample(cons, n_accept=10_000, wave=1 << 16, seed=12, cap=20_000)
    assert tal["accepted"] >= 10_000 and tal["raw"] == st.waves * (1 << 16)
    again, tal2, _ = engine.sample(cons, n_raw=tal["raw"], seed=12, cap=20_000)
    assert np.array_equal(sort_records(rec), sort_records(again))           # the set depends only on (seed, range)
    _, short, _ = engine.sample(cons, n_raw=tal["raw"] - (1 << 16), seed=12, cap=0)
    assert short["accepted"] < 10_000                                        # smallest multiple of the wave size


def test_empty_and_edge_inputs(engine):
    acc, feats = engine.filter(np.empty((0, 2), np.uint64))
    assert acc.size == 0
    rec, tal, _ = engine.sample(None, n_raw=1, seed=1)
    assert tal["raw"] == 1 and rec.shape[0] + tal["voided"] == 1
    rec, tal, st = engine.sample(None, n_raw=1000, seed=1, cap=10)           # cap smaller than the accepted count
    assert rec.shape[0] == 10 and tal["accepted"] > 10 and tal["stored"] == 10


def test_open_ended_job_watchdog(engine):
    """an accept-target job on a constraint (almost) nothing satisfies returns instead of spinning for ever"""
    cons = Constraints([None, None, seat_spec(hcp=(37, 37)), None])           # 4 hands in 6.35e11
    rec, tal, st = engine.sample(cons, n_accept=1000, wave=1 << 30, seed=1, cap=1000)
    assert tal["accepted"] < 1000 and tal["raw"] >= 1 << 36
    # and the host mirror reports it
    from bridge_mc_b200.deal import NoDealFound
    with pytest.raises(NoDealFound):
        Deal(defs=[{"hcp": [37, 37]}], engine=engine, seed=1, semantics="uniform", max_raw=1 << 30).build()


def test_auction_rules_and_first_two_calls(engine):
    """auction constraints composed from the reference's tables (config C3) and the batched class-`bidding` next"""
    rules = bidding.auction_rules(0, ["P", "P", (1, "NT"), "P", (3, "NT")])
    cons = Constraints(None, rules)
    rec, tal, _ = engine.sample(cons, n_raw=3_000_000, seed=41)
    assert 3000 < rec.shape[0] < 4600                                   # p = 1.258e-3
    opener, opening, response = bidding.first_two_calls(rec, engine)
    assert (opener == 2).all() and set(opening) == {(1, "NT")} and set(response) == {(3, "NT")}
    # on unconstrained deals: against the oracle's choose-bid, seat by seat
    raw, _, _ = engine.sample(None, n_raw=20_000, seed=42)
    opener, opening, response = bidding.first_two_calls(raw, engine)
    m = seat_masks(raw)
    op = O.choose_bid_batch(0, m.reshape(-1)).reshape(-1, 4)
    names = bidding.openings.bids()
    for i in range(0, raw.shape[0], 3):
        seats = [k for k in range(4) if op[i, k] >= 0]
        if not seats:
            assert opener[i] == -1 and opening[i] is None
            continue
        k = seats[0]
        assert opener[i] == k and opening[i] == names[op[i, k]]
        partner = (k + 2) % 4
        bid = opening[i]
        if bid == (2, "C"):
            sid, table = 3, bidding.two_c_responses
        elif bid[1] == "NT":
            sid, table = bid[0], bidding.nt_responses(bid[0])
        else:
            continue                                                    # basic-responses: checked against the Python evaluator elsewhere
        r = int(O.choose_bid_batch(sid, m[i:i + 1, partner])[0])
        assert response[i] == (None if r < 0 else table.rows[r][0])


def test_reentrancy_across_threads(engine):
    """calls arrive concurrently from several threads in the reference (hunchentoot workers, dealgen/sim threads):
    one ctx per thread, and a ctx shared by several threads, must both give the single-threaded results"""
    import threading
    from bridge_mc_b200.engine import Engine
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    want = [sort_records(engine.sample(cons, n_raw=200_000, seed=50 + i)[0]) for i in range(6)]
    got = [None] * 6
    own = [Engine(0) for _ in range(3)]

    def work(i):
        eng = own[i] if i < 3 else engine                   # threads 3..5 share the session's ctx
        for _ in range(3):
            got[i] = sort_records(eng.sample(cons, n_raw=200_000, seed=50 + i)[0])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(6):
        assert np.array_equal(got[i], want[i])
    for e in own:
        e.close()


def test_device_hist_base_matches_the_host_mirror(engine):
    """bmc_hist_base = hist-base (bridge.cl:1325-1330) over device-evaluated measure tuples: same rows, same
    first-seen order, same counts as the host mirror on rows built from the oracle's features; then the tree"""
    from bridge_mc_b200 import _lib as L
    from bridge_mc_b200._lib import shape_table
    rec, _, _ = engine.sample(None, n_raw=5_000, seed=61)          # the host mirror is O(n x distinct), like the reference
    rec = rec[np.lexsort((rec[:, 0], rec[:, 1]))]
    m = seat_masks(rec)
    f = O.features_batch(m.reshape(-1)).reshape(-1, 4, 11)
    shapes = [tuple(r) for r in shape_table().tolist()]
    measures = [(L.M_HCP, 2), (L.M_LEN, 2, 3), (L.M_LINE_HCP, 0), (L.M_TRUMP_LEN, 0, 3), (L.M_BALANCED, 2), (L.M_LOSERS, 1),
                (L.M_SHAPE, 3), (L.M_LONGEST, 0)]
    rows = []
    for i in range(rec.shape[0]):
        comb = [int(f[i, 0, s] + f[i, 2, s]) for s in range(4)]
        longest = max(range(4), key=lambda s: (comb[s], s))
        rows.append([int(f[i, 2, 8]), int(f[i, 2, 3]), int(f[i, 0, 8] + f[i, 2, 8]), comb[3], int(f[i, 2, 10]), int(f[i, 1, 9]),
                     shapes.index(tuple(sorted((int(x) for x in f[i, 3, 0:4]), reverse=True))), longest])
    for k in (1, 2, 4, 8):
        got = engine.hist_base(rec, measures[:k])
        want = hist.hist_base([r[:k] for r in rows])
        assert got == want, k
    tree = hist.histogram([r[:2] for r in rows])
    tree_dev = hist.hist_percent(hist.hist_breakdown(engine.hist_base(rec, measures[:2])))
    assert tree == tree_dev
    with pytest.raises(Exception):
        engine.hist_base(rec, measures[:4], cap=10)
    assert engine.hist_base(np.empty((0, 2), np.uint64), measures[:1]) == []


# ---- index records ----------------------------------------------------------------------------------------
def test_index_records_expand_to_the_same_deals(engine):
    """bmc_sample_indices / bmc_expand: an accepted deal shipped as the 32-bit offset of its index regenerates to exactly
    the record the 16-byte path stores -- in every dealing mode (full deal, fixed hands, seat-first, reference order)"""
    h = str2hand("♣ AJ7 ♦ AK54 ♥ 6543 ♠ 64")
    nt = [None, None, bidding.openings.form((1, "NT")), None]
    cases = [
        (lambda: None, 50_000),
        (lambda: Constraints([seat_spec(fixed=h), seat_spec(hcp=(12, 14))]), 200_000),
        (lambda: Constraints(None, nt).set_mode(1), 300_000),
        (lambda: Constraints(None, nt).set_mode(2), 1_000_003),
        (lambda: Constraints([seat_spec(hcp=(15, 17), s=(5, None)), seat_spec(hcp=(10, 11), h=(4, None))]).set_reference_order(2), 100_000),
    ]
    for make, n in cases:
        cons = make()
        first = 777_777_777_777
        rec, tal, _ = engine.sample(cons, n_raw=n, seed=33, first_index=first)
        off, tal2, st = engine.sample_indices(cons, n_raw=n, seed=33, first_index=first)
        assert off.dtype == np.uint32 and off.shape[0] == rec.shape[0] == tal["accepted"] == tal2["accepted"]
        assert len(np.unique(off)) == off.shape[0] and (off < n).all()
        for key in ("hcp", "len", "shape"):
            assert np.array_equal(tal[key], tal2[key])
        back = engine.expand(cons, off, seed=33, first_index=first)
        assert np.array_equal(sort_records(back), sort_records(rec))
    # accept-target form: offsets relative to the call's first index across several waves
    cons = Constraints(None, nt)
    off, tal, st = engine.sample_indices(cons, n_accept=50_000, wave=1 << 18, seed=34, first_index=5)
    rec, _, _ = engine.sample(cons, n_raw=int(st.raw), seed=34, first_index=5)
    assert st.raw > (1 << 18) and np.array_equal(sort_records(engine.expand(cons, off, seed=34, first_index=5)), sort_records(rec))
    assert engine.expand(cons, np.empty(0, np.uint32)).shape == (0, 2)


def test_accept_target_waves_are_cancelled_on_the_device(engine):
    """accept-target jobs enqueue their waves in batches; the waves after the one that meets the target cancel
    themselves, so the result is still `whole waves until the target is met` and raw counts only the waves that ran"""
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    wave = 1 << 16
    for target in (1, 2_000, 2_600, 30_000):
        rec, tal, st = engine.sample(cons, n_accept=target, wave=wave, seed=12, cap=40_000)
        assert tal["raw"] == st.raw == st.waves * wave and tal["accepted"] >= target
        if st.waves > 1:                                              # one wave fewer would not have been enough
            _, short, _ = engine.sample(cons, n_raw=int(st.raw) - wave, seed=12, cap=0)
            assert short["accepted"] < target
        again, _, _ = engine.sample(cons, n_raw=int(st.raw), seed=12, cap=40_000)
        assert np.array_equal(sort_records(rec), sort_records(again))


# ---- one call, several devices ----------------------------------------------------------------------------
def test_sample_multi_is_independent_of_the_device_count(engine):
    """bmc_sample_multi: contiguous index shares per device, tallies summed inside the library.  Tallies and the accepted
    set are identical for every device count (1 .. all GPUs of the box) and equal to the single-context call."""
    import torch
    from bridge_mc_b200._lib import TALLY_ALL, TALLY_SCORE
    from bridge_mc_b200.engine import sample_multi
    cs = [compscore.Contract(t).c_struct() for t in SCORE_CONTRACTS]
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    cons.set_scoring(cs, SCORE_VUL, SCORE_PAIRS)
    n = 3_000_001
    mask = TALLY_ALL | TALLY_SCORE
    rec1, tal1, _ = engine.sample(cons, n_raw=n, seed=71, first_index=11, tally_mask=mask)
    ndev = torch.cuda.device_count()
    for g in sorted({1, min(2, ndev), min(4, ndev), ndev}):
        rec, tal, st = sample_multi(list(range(g)), cons, n_raw=n, seed=71, first_index=11, tally_mask=mask, cap=200_000 * g)
        assert st.raw == n and tal["raw"] == n
        for key in ("voided", "accepted", "stored"):
            assert tal[key] == tal1[key], (g, key)
        for key in ("hcp", "len", "shape", "tricks", "imps", "imp_dropped"):
            assert np.array_equal(tal[key], tal1[key]), (g, key)
        assert np.array_equal(sort_records(rec), sort_records(rec1)), g
    # accept-target form: every device brings its share
    rec, tal, st = sample_multi(list(range(ndev)), cons, n_accept=10_000, wave=1 << 16, seed=72, cap=20_000 * ndev, tally_mask=TALLY_ALL)
    assert tal["accepted"] >= 10_000 and rec.shape[0] == tal["accepted"]
    from bridge_mc_b200._lib import BridgeMcError
    with pytest.raises(BridgeMcError):
        sample_multi([0, 0], cons, n_raw=10)


# ---- truncation must not bias the sample (Deal.build / McCase.sim) ----------------------------------------------
def test_small_requests_are_unbiased(engine):
    """A caller that takes fewer deals than a library call produced must still see an unbiased sample: the spade ace is
    with each seat a quarter of the time (sim(50) used to keep the 50 SMALLEST records of a wave)."""
    def ace_seats(deal_obj, n):
        rec = deal_obj.build_records(n)
        assert rec.shape[0] == n
        m = seat_masks(rec)
        return [int(((m[:, k] >> np.uint64(16 * 3 + 12)) & np.uint64(1)).sum()) for k in range(4)]

    for semantics in ("reference", "uniform"):
        tot = np.zeros(4)
        for seed in range(40):
            tot += ace_seats(Deal(defs=[{}], engine=engine, seed=1000 + seed, semantics=semantics), 50)
        assert tot.sum() == 2000
        assert (abs(tot - 500) < 5 * (2000 * 0.25 * 0.75) ** 0.5).all(), (semantics, tot)
    # the deals a call overdraws are handed out by the next calls, not dropped
    d = Deal(defs=[{"hcp": [15, 17]}], engine=engine, seed=5, semantics="uniform")
    a = d.build_records(10)
    used = d._next_index
    b = d.build_records(10)
    assert d._next_index == used and not np.array_equal(a, b)
    rows = McCase(Deal(engine=engine, seed=9), [lambda t, n, e, s, w: 12 in s.suits[3]]).sim(50)
    assert len(rows) == 50


# ---- the auction driver on the device (class bidding / next, bidding.cl:457-480) ---------------------------------------
def test_auction_driver_kernel(engine):
    """bmc_auction: one kernel finds the opener from any dealer, the opening, the response scheme that opening selects
    and partner's answer.  Against the oracle's choose-bid per seat (openings, nt-responses, 2C-responses), the Python
    restatement of good-opening? for basic-responses, and the reference's own example deal (bidding.cl:482-487)."""
    from oracle import rule_eval as R
    from bridge_mc_b200 import sexp
    rec, _, _ = engine.sample(None, n_raw=30_000, seed=43)
    m = seat_masks(rec)
    op = O.choose_bid_batch(0, m.reshape(-1)).reshape(-1, 4)
    f = O.features_batch(m.reshape(-1)).reshape(-1, 4, 11)
    names = bidding.openings.bids()
    tables = [bidding.response_scheme(b) for b in names]
    schemes = [t.scheme() if t is not None else None for t in tables]
    parsed = {k: [(b, sexp.read(frm)) for b, frm in t.rows] for k, t in enumerate(tables) if t is not None}
    for dealer in range(4):
        out = engine.auction(rec, bidding.openings.scheme(), schemes, dealer)
        order = [(dealer + p) % 4 for p in range(4)]
        for i in range(0, rec.shape[0], 5 if dealer else 1):
            seats = [k for k in order if op[i, k] >= 0]
            if not seats:
                assert tuple(out[i]) == (-1, -1, -1, 0)
                continue
            k = seats[0]
            row = int(op[i, k])
            assert (out["opener"][i], out["opening"][i]) == (k, row), (dealer, i)
            partner = (k + 2) % 4
            if tables[row] is None:
                assert out["response"][i] == -1 and out["status"][i] == 0
                continue
            d = {"lens": f[i, partner, 0:4].tolist(), "power": f[i, partner, 4:8].tolist(), "hcp": int(f[i, partner, 8]),
                 "losers": int(f[i, partner, 9]), "balanced": bool(f[i, partner, 10])}
            try:
                want = next((j for j, (_, frm) in enumerate(parsed[row]) if R._truth(R.evaluate(frm, d))), -1)
                assert (out["response"][i], out["status"][i]) == (want, 0), (dealer, i, names[row])
            except R.IllegalCall:
                assert out["status"][i] == 2, (dealer, i, names[row])
    assert (out["status"] == 2).any()                   # the un-spliced list of bidding.cl:249 is reached by some deals
    deal = [str2hand("N: ♠ Q1097 ♥ 3 ♦ KJ1063 ♣ K102"), str2hand("E: ♠ 62 ♥ AJ75 ♦ A942 ♣ 987"),
            str2hand("S: ♠ AKJ5 ♥ K1064 ♦ Q87 ♣ A6"), str2hand("W: ♠ 843 ♥ Q982 ♦ 5 ♣ QJ543")]
    one = engine.auction(cards.pack_deals([deal]), bidding.openings.scheme(), schemes, 0)[0]
    ref = [O.choose_bid(0, h) for h in deal]
    first = next(i for i, k in enumerate(ref) if k >= 0)
    assert (one["opener"], one["opening"]) == (first, ref[first])
    assert engine.auction(np.empty((0, 2), np.uint64), bidding.openings.scheme(), schemes).shape == (0,)


def test_play_training_setup(engine):
    """training.cl:58-79 (a22): build + longest-suit x 2 + four estimator calls + better-score, first with a stand-in
    estimator that counts the declaring line's HCP (to pin the order of the calls), then with the library's play-deal"""
    from types import SimpleNamespace
    from bridge_mc_b200 import training
    calls = []

    def fake_play_deal(trump, declarer, left, dummy, right):
        hcp = sum(max(r - 8, 0) for h in (declarer, dummy) for s in h.suits for r in s)
        calls.append(trump)
        return SimpleNamespace(tricks=hcp // 3 + (1 if trump else 0), first_lead=(0, 0))

    drill = training.play_training(fake_play_deal, training.full_random(engine, seed=77))
    deal = drill["deal"]
    assert len(deal) == 4 and sorted(c for h in deal for c in h.cards()) == list(range(52))
    ns, we = training.longest_suit(*deal), training.longest_suit(*deal[1:], deal[0])
    assert calls == [ns, None, we, None]                                  # the four candidates, in play-training's order
    best = max(range(4), key=lambda i: ([sum(max(r - 8, 0) for h in (c[0][0], c[0][2]) for s in h.suits for r in s) // 3
                                         + (1 if c[1] else 0) for c in training.drill_candidates(deal)][i], -i))
    assert (drill["hands"], drill["trump"]) == training.drill_candidates(deal)[best]
    # the real estimator (bmc_play_hands): the chosen candidate has the most declarer tricks, the first one on ties
    from bridge_mc_b200 import playdeal
    real = training.play_training(dealgen=training.full_random(engine, seed=78), engine=engine)
    outs = [playdeal.play_deal(t, *hs, engine=engine) for hs, t in training.drill_candidates(real["deal"])]
    tr = [o.score[1] for o in outs]
    assert real["outcome"].score[1] == max(tr) and (real["hands"], real["trump"]) == training.drill_candidates(real["deal"])[tr.index(max(tr))]
    assert real["lead"] == real["outcome"].tricks[0][0]
    # longest-suit: later suits win ties (bridge.cl:3032 `>=`)
    h = [str2hand("♣ AKQ2 ♦ 432 ♥ 432 ♠ 432"), None, str2hand("♣ 43 ♦ 765 ♥ 8765 ♠ 8765"), None]
    assert training.longest_suit(*h) == "S"


def test_packed_index_records_round_trip(engine):
    """bmc_sample_packed / bmc_unpack_indices: the gap-byte stream decodes to exactly the offsets bmc_sample_indices
    returns (same set, same tallies), for a sparse and a dense constraint, whole-range and accept-target jobs, small
    waves (many chunks, gaps across tiles) and the 255-escape (rare constraint: gaps of thousands)."""
    from bridge_mc_b200 import bidding
    cons_1nt = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    cons_rare = Constraints(None, [bidding.nt_responses(2).chooses((6, "NT")), None, bidding.openings.chooses((2, "NT")), None])
    for cons, kw in ((cons_1nt, dict(n_raw=3_000_017, wave=1 << 20)), (None, dict(n_raw=200_003, wave=1 << 16)),
                     (cons_rare, dict(n_raw=50_000_000, wave=1 << 24)), (cons_1nt, dict(n_accept=40_000, wave=1 << 18))):
        offs, tal_a, st_a = engine.sample_indices(cons, seed=17, first_index=12345, cap=4_000_000, **kw)
        packed, tal_b, st_b = engine.sample_packed(cons, seed=17, first_index=12345, cap_bytes=8_000_000, **kw)
        got = engine.unpack_indices(packed)
        assert st_a.accepted == st_b.accepted == got.size == offs.size and st_a.raw == st_b.raw
        assert np.array_equal(np.sort(got), np.sort(offs))
        assert tal_a["hcp"].tolist() == tal_b["hcp"].tolist()
        if cons is cons_1nt and "n_raw" in kw:
            assert packed.size < 1.15 * got.size                      # about one byte per accepted deal
        if cons is cons_rare:
            assert (packed == 255).any()                              # the skip byte is exercised
    # too small a buffer is an error with the size needed, never a silent truncation
    import ctypes as C
    from bridge_mc_b200._lib import BridgeMcError
    with pytest.raises(BridgeMcError):
        engine.sample_packed(cons_1nt, n_raw=3_000_017, seed=17, cap_bytes=50_000)
