"""The reference's exact enumerator on the device (bmc_distgen, csrc/distgen.cuh) against the reference's own golden
file, its inline KATs and the exact counts; and the sampler's exact class table (mode 4) against its CPU mirror."""
import json
import os
from math import comb

import numpy as np
import pytest

from kat_util import load_kats, unquote
from bridge_mc_b200 import bidding
from bridge_mc_b200.cards import seat_masks
from bridge_mc_b200.engine import Constraints, seat_spec
from oracle import oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0x6272696467656D63


def sort_records(rec):
    rec = np.ascontiguousarray(rec, dtype=np.uint64).reshape(-1, 2)
    return rec[np.lexsort((rec[:, 0], rec[:, 1]))]


def test_distgen_golden_file_on_the_gpu(engine):
    """bridge.cl:1047-1051 + backend/test/distgen-test.dat: the 12 012 (weight pattern) rows, in order, bit for bit --
    the reference's one golden file pins a CUDA kernel directly (hc-permuts, possible-lc-lens, permut-weight, next,
    incl. the low-card cap on open upper bounds)"""
    g = json.load(open(os.path.join(HERE, "golden", "distgen_golden.json")))
    d = engine.distgen(seat_spec(c=(5, None), d=(0, 4), h=(0, 1), hcp=(10, 15)))
    rows = [[int(w), int(p)] for w, p in zip(d["weights"], d["patterns"])]
    assert len(rows) == 12012
    assert rows == g["rows"]
    assert d["total"] == 8020911760


def test_distgen_exact_counts_and_pmfs(engine):
    """totals and pmfs from the device enumeration equal the CPU oracle's and the closed forms"""
    assert engine.distgen(None)["total"] == 635002871360                          # nil bounds: the 9-card cap quirk
    full = engine.distgen(seat_spec(c=(0, 13), d=(0, 13), h=(0, 13), s=(0, 13)))
    assert full["total"] == comb(52, 13) == 635013559600 and len(full["patterns"]) == 65536
    assert [int(x) for x in full["hcp_pmf"][:4]] == [2310789600, 5006710800, 8611542576, 15636342960] and full["hcp_pmf"][37] == 4
    for s in range(4):
        assert [int(x) for x in full["len_pmf"][s]] == [comb(13, k) * comb(39, 13 - k) for k in range(14)]
    spec = dict(hcp=(15, 17), c=(2, 5), d=(2, 5), h=(2, 4), s=(2, 4))
    box = engine.distgen(seat_spec(**spec))
    assert (len(box["patterns"]), box["total"]) == (10890, 29286389970)
    assert [int(x) for x in box["hcp_pmf"][15:18]] == [12732541524, 9607921812, 6945926634]
    o = O.Distgen(**spec)
    hcp_o, len_o = o.pmf()
    assert [int(x) for x in box["hcp_pmf"]] == hcp_o and [[int(x) for x in r] for r in box["len_pmf"]] == len_o
    assert [[int(w), int(p)] for w, p in zip(box["weights"], box["patterns"])] == [[w, p] for w, p in o.items()]


def test_distgen_on_partial_decks_and_suit_hcp_ranges(engine):
    """remaining decks (after a fixed hand / a sampled seat) and third-element ranges: rows equal the CPU oracle's"""
    from bridge_mc_b200.cards import str2hand
    rng = np.random.default_rng(5)
    h = str2hand("♣ AJ7 ♦ AK54 ♥ 6543 ♠ 64")
    full52 = (1 << 52) - 1
    cases = [(full52 & ~sum(1 << c for c in h.cards()), dict(hcp=(6, 10), h=(4, None), s=(0, 3, (1, None)))),
             (full52 & ~sum(1 << c for c in h.cards()), dict(c=(2, 4, (0, 3)), d=(0, 2))),
             (full52, dict(hcp=(20, None), s=(5, None, (5, 9))))]
    for _ in range(3):                       # random 26-card decks
        deck = sum(1 << int(c) for c in rng.choice(52, size=26, replace=False))
        cases.append((deck, dict(hcp=(5, 12), c=(1, None), s=(None, 5))))
    for deck52, spec in cases:
        lanes = int(O.mask52_to_lanes64(np.array([deck52], dtype=np.uint64))[0])
        d = engine.distgen(seat_spec(**spec), cards=lanes)
        o = O.Distgen(deck52, **spec)
        assert d["total"] == o.total
        assert [[int(w), int(p)] for w, p in zip(d["weights"], d["patterns"])] == [[w, p] for w, p in o.items()], spec
    # KATs of hc-permuts / possible-lc-lens are replayed on the oracle (tests/test_oracle_kats.py); a deck without
    # honours has no pattern at all (bridge.cl:955-957)
    low_only = sum(0x1FF << (16 * s) for s in range(4))
    d = engine.distgen(None, cards=low_only)
    assert len(d["patterns"]) == 0 and d["total"] == 0


# ---- mode 4: the exact class table ------------------------------------------------------------------------
def test_exact_mode_bit_exact_vs_cpu_mirror(engine):
    """South test-bid (1 NT): auto mode resolves to the exact seat-first sampler; the accepted set, the A1 pass count and
    the primary seat's exact acceptance (24 235 203 726 hands) equal the CPU mirror's"""
    cons = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    n = 400_000
    for first in (0, 3, 999_999_999_999):
        rec, tal, _ = engine.sample(cons, n_raw=n, seed=SEED, first_index=first)
        assert cons.mode == 4 and cons.primary == 2
        assert cons.primary_acceptance[0] == 24235203726
        x = O.ExactTable({"hcp": (15, 17), "balanced": (1, 1)})
        assert x.weight == 24235203726 and x.classes == cons.primary_acceptance[1]
        ref, flags = x.deal_batch(SEED, first, n, 2)
        good = (flags & 8 != 0) & (flags & 5 == 0)
        assert tal["raw"] == n and tal["accepted"] == int(good.sum()) == rec.shape[0]
        assert tal["voided"] == int((flags & 1 != 0).sum()) + int(((flags & 8 != 0) & (flags & 4 != 0)).sum())
        assert np.array_equal(sort_records(rec), sort_records(ref[good]))
    # every accepted South hand opens 1NT according to the oracle
    m = seat_masks(rec)
    assert (O.choose_bid_batch(0, m[:, 2]) == 2).all()


@pytest.mark.parametrize("case", ["ranges", "choose", "other_seats"])
def test_exact_mode_other_constraints(engine, case):
    """range-only seats, a choose-bid form the ranges cannot absorb (the rule program runs inside the enumeration),
    and constraints on the other seats (tested in stage B; the primary seat is never re-tested)"""
    n = 600_000
    if case == "ranges":
        cons = Constraints([None, seat_spec(hcp=(11, 14), s=(5, 6, (3, None)), h=(0, 3), losers=(5, 7)), None, None]).set_mode(4)
        x = O.ExactTable({"hcp": (11, 14), "lens": [None, None, (0, 3), (5, 6)], "power": [None, None, None, (3, 10)], "losers": (5, 7)})
        primary, check = 1, None
    elif case == "choose":
        cons = Constraints(None, [None, None, None, bidding.openings.chooses((1, "S"))]).set_mode(4)
        x = O.ExactTable(kind=2, scheme=0, k=3)
        primary, check = 3, lambda m: O.choose_bid_batch(0, m[:, 3]) == 3
    else:
        p = bidding.openings.passes()
        cons = Constraints(None, [p, p, bidding.openings.chooses((2, "NT")), None]).set_mode(4)
        x = O.ExactTable(kind=2, scheme=0, k=1)
        primary = 2
        check = lambda m: (O.choose_bid_batch(0, m[:, 2]) == 1) & (O.choose_bid_batch(0, m[:, 0]) == -1) & (O.choose_bid_batch(0, m[:, 1]) == -1)
        n = 4_000_000
    rec, tal, _ = engine.sample(cons, n_raw=n, seed=77, first_index=10)
    assert cons.mode == 4 and cons.primary == primary and cons.primary_acceptance == (x.weight, x.classes)
    ref, flags = x.deal_batch(77, 10, n, primary)
    cand = ref[(flags & 8 != 0) & (flags & 5 == 0)]
    if case == "other_seats":
        cand = cand[check(seat_masks(cand))]
    assert rec.shape[0] == tal["accepted"] > 50
    assert np.array_equal(sort_records(rec), sort_records(cand))
    if check is not None:
        assert check(seat_masks(rec)).all()


def test_exact_mode_distribution_and_jit(engine):
    """accepted deals are uniform over the valid deals whatever the mode: tallies of modes 1, 2 and 4 agree statistically
    (same constraint, 2^26 raw deals each), interpreter and NVRTC builds of mode 4 agree bit for bit"""
    from scipy.stats import chi2
    p = bidding.openings.passes()
    nt = bidding.nt_responses(1)
    rules = [f"(and {p} {nt.chooses((3, 'NT'))})", p, bidding.openings.chooses((1, "NT")), p]
    tal = {}
    for mode in (1, 2, 4):
        cons = Constraints(None, rules).set_mode(mode).set_jit(0)
        _, tal[mode], _ = engine.sample(cons, n_raw=1 << 26, seed=91, cap=0)
    for mode in (2, 4):
        for k in range(4):
            a, b = tal[1]["hcp"][k].astype(float), tal[mode]["hcp"][k].astype(float)
            keep = (a + b) >= 10
            a, b = a[keep], b[keep]
            ea = (a + b) * a.sum() / (a.sum() + b.sum())
            eb = (a + b) * b.sum() / (a.sum() + b.sum())
            stat = ((a - ea) ** 2 / ea + (b - eb) ** 2 / eb).sum()
            assert chi2.sf(stat, keep.sum() - 1) > 1e-4, (mode, k)
        assert abs(tal[mode]["accepted"] - tal[1]["accepted"]) < 6 * (2 * tal[1]["accepted"]) ** 0.5
    a = Constraints(None, rules).set_mode(4).set_jit(0)
    b = Constraints(None, rules).set_mode(4).set_jit(1)
    ra, ta, _ = engine.sample(a, n_raw=3_000_000, seed=92)
    rb, tb, _ = engine.sample(b, n_raw=3_000_000, seed=92)
    assert b.jit_ms > 0 and ta["accepted"] == tb["accepted"] > 1000
    assert np.array_equal(sort_records(ra), sort_records(rb))
    for key in ("hcp", "len", "shape"):
        assert np.array_equal(ta[key], tb[key])


def test_a_seat_nothing_satisfies_is_reported(engine):
    from bridge_mc_b200._lib import BadRange
    # 37 HCP needs A K Q in every suit and a jack: 4-3-3-3, never five clubs -- but no single range says so
    cons = Constraints(None, [None, None, "(and (> hcp 36) (> C 4))", None])
    with pytest.raises(BadRange):
        engine.sample(cons, n_raw=1000)
    # mode 1 does not enumerate: it simply accepts nothing
    _, tal, _ = engine.sample(Constraints(None, [None, None, "(and (> hcp 36) (> C 4))", None]).set_mode(1), n_raw=100_000, cap=0)
    assert tal["accepted"] == 0
