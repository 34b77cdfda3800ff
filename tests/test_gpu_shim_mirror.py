"""lisp/bridgemc-cffi.lisp cannot be executed in the build image (no Lisp implementation).  This file makes, through
raw ctypes, the SAME foreign calls with the SAME argument conventions the shim makes for each re-pointed entry point --
build / refill, sim, assess-hand, test-bid, choose-bid, contract-score, base-score, match-points, tallies->base,
records-hist-base, sample-score-tallies, the multi-device path -- and checks the results against the reference's KATs and
the CPU oracle.  (The object wrappers of bridge_mc_b200/engine.py are deliberately NOT used here.)"""
import ctypes as C

import numpy as np
import pytest

from kat_util import bid_of, hand_of, load_kats, truth, unquote
from bridge_mc_b200 import _lib, bidding, compscore, hist
from bridge_mc_b200._lib import Contract, SeatSpec, Stats
from bridge_mc_b200.cards import Hand, seat_masks, str2hand, unpack_deal
from oracle import oracle as O

pytestmark = pytest.mark.gpu
NULL = None


@pytest.fixture(scope="module")
def ctx():
    L = _lib.load()
    h = C.c_void_p()
    assert L.bmc_init(0, C.byref(h)) == 0                    # (bmc-init (first *bmc-devices*) out)
    yield L, h
    L.bmc_shutdown(h)


def hand_to_record(hand: Hand):
    """the shim's hand->record: seat 0 = the hand, the other 39 cards to seats 1..3 in card order"""
    mine = hand.mask64()
    lo = hi = n = 0
    for suit in range(4):
        for rank in range(13):
            bit = 16 * suit + rank
            if not (mine >> bit) & 1:
                seat = 1 + n // 13
                lo |= (seat & 1) << bit
                hi |= (seat >> 1) << bit
                n += 1
    return np.array([[lo, hi]], dtype=np.uint64)


def filter_hand(ctx, hand, table_text=None):
    """the shim's filter-hand: bmc_filter(ctx, NULL, scheme-or-NULL, rec, 1, NULL, feats) -> seat 0"""
    L, h = ctx
    rec = hand_to_record(hand)
    feats = np.zeros((1, 4), dtype=_lib.FEATURE_DTYPE)
    scheme = C.c_void_p()
    if table_text is not None:
        assert L.bmc_scheme_create(table_text.encode(), C.byref(scheme)) == 0, L.bmc_last_error()
    assert L.bmc_filter(h, NULL, scheme if table_text is not None else NULL, rec.ctypes.data, 1, NULL, feats.ctypes.data) == 0
    if table_text is not None:
        L.bmc_scheme_destroy(scheme)
    f = feats[0, 0]
    return [int(x) for x in f["len"]], [int(x) for x in f["suit_hcp"]], int(f["losers"]), int(f["hcp"]), bool(f["balanced"]), int(f["bid"])


def test_assess_hand_test_bid_choose_bid_as_the_shim_calls_them(ctx):
    """bidding.cl KATs through filter-hand: assess-hand #'balanced? (6), test-bid (28: a ONE-ROW table with the bid's
    rule, so only that rule is evaluated), choose-bid over openings / nt-responses / 2C-responses (31)"""
    n = 0
    for where, form, exp in load_kats("ASSESS-HAND"):
        assert filter_hand(ctx, hand_of(form[1]))[4] == truth(exp), where
        lens, power, losers, hcp, bal, _ = filter_hand(ctx, hand_of(form[1]))
        ref = O.assess(hand_of(form[1]))
        assert (lens, power, losers) == (ref["lens"], ref["power"], ref["losers"])          # (lens power loosers)
        n += 1
    for where, form, exp in load_kats("TEST-BID"):
        bid = bid_of(form[1])
        one_row = f"((({bid[0]} {bid[1]}) . {bidding.openings.form(bid)}))"
        assert (filter_hand(ctx, hand_of(form[2]), one_row)[5] == 0) == truth(exp), where
        n += 1
    for where, form, exp in load_kats("CHOOSE-BID"):
        if len(form) == 2:
            table = bidding.openings
        elif form[2][0] == "NT-RESPONSES":
            table = bidding.nt_responses(form[2][1])
        elif form[2] == "2C-RESPONSES":
            table = bidding.two_c_responses
        else:
            continue
        idx = filter_hand(ctx, hand_of(form[1]), table.sexpr())[5]
        got = None if idx < 0 else table.rows[idx][0]
        assert got == bid_of(exp), where
        n += 1
    assert n == 6 + 28 + 31
    # an illegal form that evaluation reaches: bid = -2 (the shim signals the reference's error)
    t = bidding.basic_responses((1, "H"))
    assert filter_hand(ctx, str2hand("♣ 72 ♦ K854 ♥ K653 ♠ AQ8"), t.sexpr())[5] == -2


def test_scoring_entry_points_as_the_shim_calls_them(ctx):
    """contract-score / base-score: bmc_score with ONE contract, one trick count; match-points: bmc_match_points on
    one pair; compscore.cl KATs"""
    L, h = ctx

    def score_1(level, suit, declarer, dbl, tricks, vul):
        c = (Contract * 1)(Contract(level, suit, declarer, dbl))
        v = (C.c_uint8 * 1)(1 if vul else 0)
        tr = (C.c_uint8 * 1)(tricks)
        out = (C.c_int32 * 1)()
        assert L.bmc_score(h, c, v, 1, tr, 1, out) == 0
        return out[0]

    assert score_1(0, 2, -1, 0, 10, False) == 420                       # (contract-score 'h 10)
    assert score_1(3, 4, -1, 0, 9, True) == 600                         # (contract-score nil 9 :vulnerable t :level 3)
    for where, form, exp in load_kats("BASE-SCORE"):
        c = compscore.Contract(str(form[1][3])).c_struct()
        tricks = compscore.Contract(str(form[1][3])).tricks()
        vul = len(form) > 2 and truth(form[3])
        assert score_1(c.level, c.suit, c.declarer, c.dbl, tricks, vul) == exp, where
    for where, form, exp in load_kats("MATCH-POINTS"):
        a, b = compscore.Contract(str(form[1][3])), compscore.Contract(str(form[2][3]))
        vul = str(unquote(form[4])).lower() if len(form) > 3 else None
        sa = score_1(*[getattr(a.c_struct(), k) for k in ("level", "suit", "declarer", "dbl")], a.tricks(), vul in ("ns", "both"))
        sb = score_1(*[getattr(b.c_struct(), k) for k in ("level", "suit", "declarer", "dbl")], b.tricks(), vul in ("ew", "both"))
        pa, pb = (C.c_int32 * 1)(sa), (C.c_int32 * 1)(sb)
        imps, status = (C.c_int8 * 1)(), (C.c_uint8 * 1)()
        assert L.bmc_match_points(h, pa, pb, 1, imps, status) == 0
        assert status[0] == 0 and imps[0] == exp, where
    pa, pb = (C.c_int32 * 1)(7600), (C.c_int32 * 1)(-7600)
    imps, status = (C.c_int8 * 1)(), (C.c_uint8 * 1)()
    assert L.bmc_match_points(h, pa, pb, 1, imps, status) == 0 and status[0] == 1      # the shim signals a type-error


def make_deal_handle(L, const_hands, defs):
    """the shim's make-deal-handle: specs cleared to -1, fixed hands first, then the :~ defs; reference order"""
    specs = (SeatSpec * 4)()
    for k in range(4):
        C.memset(C.byref(specs[k]), 0, C.sizeof(SeatSpec))
        specs[k].hcp[0] = specs[k].hcp[1] = specs[k].losers[0] = specs[k].losers[1] = -1
        for s in range(4):
            specs[k].len[s][0] = specs[k].len[s][1] = specs[k].suit_hcp[s][0] = specs[k].suit_hcp[s][1] = -1
    for k, hnd in enumerate(const_hands):
        specs[k].fixed = 1
        specs[k].fixed_mask = hnd.mask64()
    for k, d in enumerate(defs, start=len(const_hands)):
        def fill(arr, spec):
            if spec is None:
                return
            arr[0] = -1 if spec[0] is None else spec[0]
            arr[1] = -1 if len(spec) < 2 or spec[1] is None else spec[1]
        fill(specs[k].hcp, d.get("hcp"))
        for i, key in enumerate("cdhs"):
            fill(specs[k].len[i], d.get(key))
            fill(specs[k].suit_hcp[i], d[key][2] if d.get(key) is not None and len(d[key]) > 2 else None)
    handle = C.c_void_p()
    assert L.bmc_constraints_create(specs, NULL, C.byref(handle)) == 0, L.bmc_last_error()
    assert L.bmc_constraints_set_reference_order(handle, max(1, len(defs))) == 0
    return handle


def sample_records(ctx, handle, seed, first_index, n, tally_mask=0, tallies=None, devices=None):
    """the shim's sample-records: bmc_sample(ctx, handle, seed, first, 0, n, 16384, mask, recs, n + 16384, tallies-or-NULL,
    stats) -- or bmc_sample_multi over *bmc-devices*; the next job continues at first + stats.raw"""
    L, h = ctx
    cap = n + 16384
    recs = np.zeros((cap, 2), dtype=np.uint64)
    st = Stats()
    if devices and len(devices) > 1:
        devs = (C.c_int * len(devices))(*devices)
        rc = L.bmc_sample_multi(devs, len(devices), handle, seed, first_index, 0, n, 16384, tally_mask, recs.ctypes.data, cap,
                                tallies, C.byref(st))
    else:
        rc = L.bmc_sample(h, handle, seed, first_index, 0, n, 16384, tally_mask, recs.ctypes.data, cap, tallies, C.byref(st))
    assert rc == 0, L.bmc_last_error()
    return recs[:st.stored], st


def test_build_and_sim_as_the_shim_calls_them(ctx):
    """`build` / `sim` of a deal with a fixed South hand and a ranged North (rest.cl:264-277 shape): one bmc_sample per
    refill with n_accept = runs; rows (hands m1 ...) then go through the reference's own hist-base pipeline"""
    L, h = ctx
    south = str2hand("S: ♠ AKJ5 ♥ K1064 ♦ Q87 ♣ A6")
    handle = make_deal_handle(L, [south], [{"hcp": [10, 12], "s": [4, None]}])
    recs, st = sample_records(ctx, handle, 12345, 0, 500)
    assert recs.shape[0] >= 500 and st.raw >= 500 and st.raw % 16384 == 0        # whole waves; raw = indices used
    more, st2 = sample_records(ctx, handle, 12345, int(st.raw), 500)            # (incf next-index raw)
    assert not np.array_equal(recs[:500], more[:500])
    rows = []
    for rec in recs[:500]:
        hands = unpack_deal(rec, ids=None)                                        # record->hands
        assert hands[0] == south
        hcp = sum(max(r - 8, 0) for s in hands[1].suits for r in s)
        assert 10 <= hcp <= 12 and len(hands[1].suits[3]) >= 4
        rows.append([len(hands[0].suits[3]) + len(hands[1].suits[3])])           # a measure: (trump n e s w) -> value
    tree = hist.histogram(rows)
    assert sum(e[1] for e in tree) == 1
    L.bmc_constraints_destroy(handle)                                             # the deal's finaliser


def test_tallies_as_hist_base_and_device_hist_base(ctx):
    """tallies->base: the tally words of a tally-only bmc_sample as hist-base's BASE rows; records-hist-base:
    bmc_hist_base with bmc_measure structs; both equal the host hist-base over the same deals"""
    L, h = ctx
    handle = make_deal_handle(L, [], [{"hcp": [15, 17], "c": [2, 5], "d": [2, 5], "h": [2, 4], "s": [2, 4]}])
    tallies = (C.c_uint64 * _lib.TALLY_WORDS)()
    st = Stats()
    n = 20000
    assert L.bmc_sample(h, handle, 7, 0, n, 0, 0, 7, NULL, 0, C.cast(tallies, C.POINTER(_lib.Tallies)), C.byref(st)) == 0     # sample-tallies
    words = np.ctypeslib.as_array(tallies)
    base = [[v, int(words[4 + 40 * 0 + v])] for v in range(38) if words[4 + 40 * 0 + v]]      # (tallies->base t :hcp 0)
    assert sum(c for _, c in base) == n and {v for v, _ in base} == {15, 16, 17}
    recs = np.zeros((n, 2), dtype=np.uint64)
    assert L.bmc_sample(h, handle, 7, 0, n, 0, 0, 0, recs.ctypes.data, n, NULL, C.byref(st)) == 0
    f = O.features_batch(seat_masks(recs)[:, 0])
    assert sorted(hist.hist_base([int(x) for x in f[:, 8]])) == sorted(base)
    assert hist.hist_base([15, 40], base)[-1] == [40, 1]                          # extended like the reference's KAT
    # records-hist-base: k = 2 measures (HCP of seat 0, its spade length)
    ms = (_lib.Measure * 2)(_lib.Measure(0, 0, 0, 0), _lib.Measure(1, 0, 3, 0))
    cap = 65536
    tuples = np.zeros((cap, 2), dtype=np.uint8)
    counts = np.zeros(cap, dtype=np.uint64)
    nu = C.c_uint64()
    sub = np.ascontiguousarray(recs[:3000])
    assert L.bmc_hist_base(h, sub.ctypes.data, 3000, ms, 2, tuples.ctypes.data, counts.ctypes.data, cap, C.byref(nu)) == 0
    got = [[[int(x) for x in tuples[i]], int(counts[i])] for i in range(nu.value)]
    assert got == hist.hist_base([[int(r[8]), int(r[3])] for r in f[:3000]])
    L.bmc_constraints_destroy(handle)


def test_score_tallies_and_multi_device_as_the_shim_calls_them(ctx):
    """sample-score-tallies: bmc_constraints_set_scoring + bmc_sample with tally_mask = 8 and records = NULL; the
    multi-device branch of sample-records (bmc_sample_multi with NULL tallies)"""
    import torch
    L, h = ctx
    handle = make_deal_handle(L, [], [{"hcp": [15, 17], "c": [2, 5], "d": [2, 5], "h": [2, 4], "s": [2, 4]}])
    names = ("1NTN", "3NTN", "4HN", "4SN")
    cs = (Contract * 4)(*[compscore.Contract(t).c_struct() for t in names])
    vul = (C.c_uint8 * 4)(0, 1, 0, 1)
    pr = (C.c_uint8 * 4)(0, 1, 2, 3)
    assert L.bmc_constraints_set_scoring(handle, cs, vul, 4, pr, 2) == 0
    tallies = (C.c_uint64 * _lib.TALLY_WORDS)()
    st = Stats()
    assert L.bmc_sample(h, handle, 9, 0, 0, 30000, 16384, 8, NULL, 0, C.cast(tallies, C.POINTER(_lib.Tallies)), C.byref(st)) == 0
    words = np.ctypeslib.as_array(tallies).astype(np.int64)
    tricks = words[548:548 + 112].reshape(8, 14)
    imps = words[548 + 112:548 + 112 + 392].reshape(8, 49)
    recs = np.zeros((int(st.raw), 2), dtype=np.uint64)
    assert L.bmc_sample(h, handle, 9, 0, int(st.raw), 0, 0, 0, recs.ctypes.data, recs.shape[0], NULL, C.byref(st)) == 0
    t_o, i_o, d_o = O.score_tally_records(9, recs[:st.stored], list(cs), [0, 1, 0, 1], [0, 1, 2, 3])
    assert np.array_equal(tricks, t_o) and np.array_equal(imps, i_o) and words[548 + 504:].tolist() == d_o.tolist()
    rows = [[v - 24, int(c)] for v, c in enumerate(imps[0]) if c]                   # ((imps count) ...) for hist-breakdown
    assert sum(c for _, c in rows) + int(d_o[0]) == st.stored
    devices = list(range(torch.cuda.device_count()))
    got, st = sample_records(ctx, handle, 5, 0, 1000, devices=devices)
    assert got.shape[0] >= 1000
    L.bmc_constraints_destroy(handle)


def test_play_deal_as_the_shim_calls_it(ctx):
    """lisp/bridgemc-cffi.lisp `play-deal` / `best-tricks` / `records-best-tricks`: bmc_play_hands on four lane masks with
    one int8 trump, the 72-byte result decoded field by field; bmc_best_tricks on a record buffer with scalar trump /
    declarer.  Checked against the reference's own run (tests/golden/playdeal_ref.json)."""
    import json
    import os
    L, h = ctx
    rows = [r for r in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "playdeal_ref.json")))["rows"] if "tricks" in r]
    for r in rows[:6] + rows[-6:]:
        hands = (C.c_uint64 * 4)(*[sum(sum(1 << x for x in s) << (16 * i) for i, s in enumerate(hd)) for hd in r["hands"]])
        tr = C.c_int8(-1 if r["trump"] is None else r["trump"])
        out = (C.c_uint8 * 72)()
        assert L.bmc_play_hands(h, hands, 1, C.byref(tr), 0, 0, out) == 0
        status, n = out[0], out[1]
        assert status == 0 and n == len(r["tricks"]) and [out[2], out[3]] == r["score"]
        cards = [[None if out[4 + 4 * t + i] == 255 else [out[4 + 4 * t + i] >> 4, out[4 + 4 * t + i] & 15] for i in range(4)] for t in range(n)]
        assert cards == r["tricks"] and [out[56 + t] for t in range(n)] == r["winner_nos"]
    full = [r for r in rows if sum(len(s) for s in r["hands"][0]) == 13][:4]
    rec = (C.c_uint64 * (2 * len(full)))()
    for i, r in enumerate(full):                                   # declarer = seat 0: N E S W = declarer, left, dummy, right
        masks = [sum(sum(1 << x for x in s) << (16 * k) for k, s in enumerate(hd)) for hd in r["hands"]]
        rec[2 * i] = masks[1] | masks[3]
        rec[2 * i + 1] = masks[2] | masks[3]
    tr, de = C.c_int8(-1), C.c_uint8(0)
    tricks = (C.c_uint8 * len(full))()
    nt = [r for r in full]
    assert L.bmc_best_tricks(h, rec, len(full), C.byref(tr), C.byref(de), 0, 0, tricks) == 0
    for i, r in enumerate(nt):
        if r["trump"] is None:
            assert tricks[i] == r["score"][1]


def test_packed_offsets_as_the_shim_calls_them(ctx):
    """lisp/bridgemc-cffi.lisp `sample-packed-offsets`: bmc_sample_packed with n_raw 0, an accept target, waves of 262 144,
    tally mask 0, buffers sized for the target plus one whole wave, then bmc_unpack_indices; the offsets expand to deals
    that satisfy the constraint."""
    L, h = ctx
    specs = (SeatSpec * 4)()
    for k in range(4):
        for f in ("hcp", "losers"):
            getattr(specs[k], f)[0] = getattr(specs[k], f)[1] = -1
        for s in range(4):
            specs[k].len[s][0] = specs[k].len[s][1] = specs[k].suit_hcp[s][0] = specs[k].suit_hcp[s][1] = -1
    specs[2].hcp[0], specs[2].hcp[1] = 20, 22
    cons = C.c_void_p()
    assert L.bmc_constraints_create(specs, None, C.byref(cons)) == 0
    n_accept, wave = 5000, 262144
    cap = n_accept + wave
    cap_bytes = 2 * cap + (1 << 16)
    buf = (C.c_uint8 * cap_bytes)()
    nb, n = C.c_uint64(), C.c_uint64()
    tallies, stats = _lib.Tallies(), Stats()
    assert L.bmc_sample_packed(h, cons, 7, 1000, 0, n_accept, wave, 0, buf, cap_bytes, C.byref(nb), C.byref(tallies), C.byref(stats)) == 0
    offs = (C.c_uint32 * cap)()
    assert L.bmc_unpack_indices(buf, nb.value, offs, cap, C.byref(n)) == 0
    assert n.value == stats.accepted >= n_accept
    rec = (C.c_uint64 * (2 * n.value))()
    assert L.bmc_expand(h, cons, 7, 1000, offs, n.value, rec) == 0
    records = np.frombuffer(rec, dtype=np.uint64).reshape(-1, 2)
    f = O.features_batch(seat_masks(records)[:, 2].copy()).reshape(-1, 11)
    assert ((f[:, 8] >= 20) & (f[:, 8] <= 22)).all()
    L.bmc_constraints_destroy(cons)
