"""The CPU oracle (oracle/bmc_oracle.c) against every golden vector the reference
holds for the hot path: its inline known-answer tests and distgen-test.dat.
This is what pins the oracle (no Lisp implementation exists in this image)."""
import json
import os
from math import comb

import pytest

from kat_util import bid_of, hand_of, load_kats, truth, unquote
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))

OPENINGS = [(2, "C"), (2, "NT"), (1, "NT"), (1, "S"), (1, "H"), (1, "D"), (1, "C"), (2, "D"), (2, "H"), (2, "S"),
            (3, "C"), (3, "D"), (3, "H"), (3, "S")]
NT1 = [(2, "C"), (2, "D"), (2, "H"), (2, "S"), (3, "C"), (2, "NT"), (3, "NT"), (4, "NT"), (6, "NT")]
NT2 = [(3, "C"), (3, "D"), (3, "H"), (3, "S"), (3, "NT"), (4, "NT"), (6, "NT")]
TWO_C = [(2, "D"), (2, "H"), (2, "S"), (2, "NT"), (3, "C"), (3, "D")]


def test_suit_loosers_kats():
    kats = load_kats("SUIT-LOOSERS")
    assert len(kats) == 7
    for where, form, exp in kats:
        ranks = unquote(form[1]) or []
        assert O.lib().orc_suit_losers(sum(1 << r for r in ranks)) == exp, where


def test_balanced_kats():
    kats = [k for k in load_kats("ASSESS-HAND")]
    assert len(kats) == 6
    for where, form, exp in kats:
        assert form[2] == ["FUNCTION", "BALANCED?"]
        assert O.assess(hand_of(form[1]))["balanced"] == truth(exp), where


def test_test_bid_kats():
    kats = load_kats("TEST-BID")
    assert len(kats) == 28
    for where, form, exp in kats:
        k = OPENINGS.index(bid_of(form[1]))
        assert O.test_bid(0, k, hand_of(form[2])) == truth(exp), where


def _scheme_of(form):
    if len(form) == 2:
        return 0, OPENINGS
    s = form[2]
    if s == ["NT-RESPONSES", 1]:
        return 1, NT1
    if s == ["NT-RESPONSES", 2]:
        return 2, NT2
    if s == "2C-RESPONSES":
        return 3, TWO_C
    return None, None


def test_choose_bid_kats():
    n = 0
    for where, form, exp in load_kats("CHOOSE-BID"):
        sid, table = _scheme_of(form)
        if sid is None:
            continue            # basic-responses / further-bid tables: generated on the host, checked on the GPU path
        k = O.choose_bid(sid, hand_of(form[1]))
        got = None if k < 0 else table[k]
        assert got == bid_of(exp), where
        n += 1
    assert n == 10 + 13 + 8


SUITS = {"C": 0, "D": 1, "H": 2, "S": 3, "NT": 4, "♣": 0, "♦": 1, "♥": 2, "♠": 3}


def _parse_contract(text):
    """independent mini-parser for the KAT strings: <level><suit>[decl][x|xx][=|+n|-n]"""
    import re
    m = re.fullmatch(r"([1-7])(NT|[CDHS♣♦♥♠])([NESW]?)(xx|x)?(=|[+-]\d+)?", text)
    assert m, text
    level, suit, decl, dbl, out = m.groups()
    tricks = int(level) + 6 + (0 if out in (None, "=") else int(out))
    return int(level), SUITS[suit], "NESW".index(decl) if decl else -1, {None: 0, "x": 1, "xx": 2}[dbl], tricks


def _contract_arg(form):
    assert form[0] == "MAKE-INSTANCE" and form[2] == ":="
    return _parse_contract(str(form[3]))


def test_base_score_kats():
    kats = load_kats("BASE-SCORE")
    assert len(kats) == 7
    for where, form, exp in kats:
        level, suit, decl, dbl, tricks = _contract_arg(form[1])
        vul = 1 if len(form) > 2 and form[2] == ":VULNERABLE" and truth(form[3]) else 0
        assert O.lib().orc_base_score(level, suit, decl, dbl, tricks, vul) == exp, where


def test_match_points_kats():
    kats = load_kats("MATCH-POINTS")
    assert len(kats) == 3
    for where, form, exp in kats:
        a, b = _contract_arg(form[1]), _contract_arg(form[2])
        vul = str(unquote(form[4])).lower() if len(form) > 3 else None
        sa = O.lib().orc_base_score(a[0], a[1], a[2], a[3], a[4], 1 if vul in ("ns", "both") else 0)
        sb = O.lib().orc_base_score(b[0], b[1], b[2], b[3], b[4], 1 if vul in ("ew", "both") else 0)
        assert O.imps(sa - sb) == exp, where


def test_contract_score_properties():
    """unpinned branches follow the source text: spot values computed by hand from bridge.cl:3045-3069"""
    cs = O.lib().orc_contract_score
    assert cs(4, 9, 0, 0, 3) == 400            # 3NT=
    assert cs(4, 13, 1, 0, 7) == 220 + 2000    # 7NT= vulnerable
    assert cs(2, 12, 1, 0, 6) == 180 + 1250    # 6H= vulnerable
    assert cs(3, 10, 0, 2, 4) == 480 + 300 + 100   # 4Sxx=
    assert cs(3, 11, 0, 2, 4) == 480 + 300 + 100 + 100   # redoubled overtrick at the doubled rate (kept quirk)
    assert cs(0, 6, 0, 1, 5) == -800 + 300 * (-5 + 3)    # 5Cx-5 not vulnerable: -1400
    assert cs(0, 7, 0, 1, 5) == -1100                    # -4 doubled not vulnerable (non-standard, kept)
    assert cs(1, 8, 1, 0, 4) == -200
    assert cs(1, 9, 0, 0, 0) == 110                      # level defaults to tricks-6
    assert cs(1, 5, 0, 0, 0) == -100                     # level defaults to 1


def test_combinations_and_permut_weight_kats():
    for where, form, exp in load_kats("COMBINATIONS"):
        assert O.lib().orc_combinations(form[1], form[2]) == exp, where
    L = O.lib()
    for where, form, exp in load_kats("PERMUT-WEIGHT"):
        supply, permut = unquote(form[1]), unquote(form[2])
        supply = list(supply) + [0] * (len(permut) - len(supply))      # add-zeros
        w = 1
        for s, p in zip(supply, permut):
            w *= L.orc_combinations(s, p)
        assert w == exp, where


def _supply_mask(supply):
    """reference supply rows (n_low nJ nQ nK nA) -> 52-bit card mask with the low cards taken from rank 0 up"""
    m = 0
    for s, row in enumerate(supply):
        for r in range(row[0]):
            m |= 1 << (13 * s + r)
        for h in range(4):
            if row[1 + h]:
                m |= 1 << (13 * s + 9 + h)
    return m


def _pattern_rows(p16, nsuits):
    return [[(p16 >> (15 - (4 * s + h))) & 1 for h in range(4)] for s in range(nsuits)]


def test_hc_permuts_kats():
    kats = load_kats("HC-PERMUTS")
    assert len(kats) == 4
    for where, form, exp in kats:
        supply = unquote(form[1])
        kw = {}
        rest = form[2:]
        for i in range(0, len(rest), 2):
            kw[str(rest[i])] = unquote(rest[i + 1])
        spec = {}
        if ":HCP" in kw:
            spec["hcp"] = tuple(kw[":HCP"])
        if ":SUITDEF" in kw:
            for name, sd in zip("cdhs", kw[":SUITDEF"]):
                if sd is not None:
                    spec[name] = (sd[0], sd[1], tuple(sd[2]) if isinstance(sd[2], list) else sd[2])
        # hc-permuts ignores suit lengths: give every suit an explicit (0 13) so no pattern has zero weight rows removed
        g = O.Distgen(_supply_mask(supply), **spec)
        got = [_pattern_rows(p, len(supply)) for _, p in g.items()]
        assert got == (exp or []), where
    where, form, exp = load_kats("LENGTH")[0]
    assert form[1][0] == "HC-PERMUTS"
    assert O.Distgen(_supply_mask(unquote(form[1][1]))).npat == exp == 1024


def test_possible_lc_lens_kats():
    kats = load_kats("POSSIBLE-LC-LENS")
    assert len(kats) == 2
    for where, form, exp in kats:
        pattern, lc_supply, lens = unquote(form[1]), unquote(form[2]), unquote(form[3])
        # build a deck with that low-card supply and every honour, then look the pattern up
        mask = _supply_mask([[n, 1, 1, 1, 1] for n in lc_supply])
        spec = {k: (v if isinstance(v, int) else tuple(v)) for k, v in zip("cdhs", lens)}
        g = O.Distgen(mask, **spec)
        p16 = 0
        for b in [x for suit in pattern for x in suit]:
            p16 = (p16 << 1) | b
        w = dict((p, wt) for wt, p in g.items())[p16]
        expect = sum(comb(lc_supply[0], t[0]) * comb(lc_supply[1], t[1]) * comb(lc_supply[2], t[2]) * comb(lc_supply[3], t[3])
                     for t in (exp or []))
        assert w == expect, where


def test_distgen_golden_file():
    """bridge.cl:1047-1051 + backend/test/distgen-test.dat: 12 012 (weight pattern) rows in order"""
    g = json.load(open(os.path.join(HERE, "golden", "distgen_golden.json")))
    d = O.Distgen(c=(5, None), d=(0, 4), h=(0, 1), hcp=(10, 15))
    rows = [[w, p] for w, p in d.items()]
    assert len(rows) == 12012
    assert rows == g["rows"]
    assert d.total == 8020911760


def test_exact_counts():
    """SURVEY appendix B facts (all from the enumeration)"""
    assert O.Distgen().total == 635002871360                       # nil bounds: the 9-card cap quirk
    full = O.Distgen(c=(0, 13), d=(0, 13), h=(0, 13), s=(0, 13))
    assert full.total == comb(52, 13) == 635013559600
    box = O.Distgen(hcp=(15, 17), c=(2, 5), d=(2, 5), h=(2, 4), s=(2, 4))
    assert (box.npat, box.total) == (10890, 29286389970)
    hcp, lens = box.pmf()
    assert hcp[15:18] == [12732541524, 9607921812, 6945926634]
    hcp, lens = full.pmf()
    assert hcp[:4] == [2310789600, 5006710800, 8611542576, 15636342960] and hcp[37] == 4
    for s in range(4):
        assert lens[s] == [comb(13, k) * comb(39, 13 - k) for k in range(14)]


def test_reference_sampler_respects_ranges():
    """`build` with one ranged seat: every sampled hand is inside the ranges (bridge.cl:2989-3003 style)"""
    g = O.Distgen(hcp=(15, 17), c=(2, 5), d=(2, 5), h=(2, 4), s=(2, 4))
    deals = g.build_batch(300, seed=5)
    f = O.features_batch(O.mask52_to_lanes64(deals[:, 0]))
    assert ((f[:, 8] >= 15) & (f[:, 8] <= 17)).all()
    assert (f[:, 0:2] >= 2).all() and (f[:, 0:2] <= 5).all() and (f[:, 2:4] >= 2).all() and (f[:, 2:4] <= 4).all()
    allc = deals[:, 0] | deals[:, 1] | deals[:, 2] | deals[:, 3]
    assert (allc == (1 << 52) - 1).all()


def test_gpu_deal_mirror_is_a_deal_and_void_rate():
    rec, void = O.gpu_deal_batch(12345, 0, 20000)
    from bridge_mc_b200.cards import seat_masks
    m = seat_masks(rec)
    for k in range(4):
        assert all(bin(int(x)).count("1") == 13 for x in m[:500, k])
    assert 0.0005 < void.mean() < 0.004


def test_best_effort_cpu_comparator_is_a_valid_sampler():
    """bench.py's second CPU comparator (not the reference algorithm): its acceptance of test-bid (1 NT) must
    agree with the exact probability 24 235 203 726 / C(52,13), and it is deterministic in (seed, index)."""
    import ctypes as C
    L = O.lib()
    n = 1_000_000
    c1, c2 = C.c_uint64(), C.c_uint64()
    a1 = L.orc_cpu_best_run(5, 1 << 33, n, C.byref(c1))
    a2 = L.orc_cpu_best_run(5, 1 << 33, n, C.byref(c2))
    assert (a1, c1.value) == (a2, c2.value)
    p = 24235203726 / 635013559600
    assert abs(a1 - n * p) < 4 * (n * p * (1 - p)) ** 0.5


def test_randcar_weights_statistical():
    """bridge.cl:298-303: 1001 repetitions of 161 weighted draws from ((1 1) (5 2) (10 3)): every value <= 3 and
    every sum in (380, 440) -- a 3.5-sigma window on the upper side, so the reference's own run fails about one
    time in five; here the generator is seeded, the outcome is fixed"""
    import ctypes as C
    L = O.lib()
    L.orc_randcar_weights.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint64)]
    w = (C.c_uint64 * 3)(1, 5, 10)
    st = C.c_uint64(20261020)
    counts = [0, 0, 0]
    for _ in range(1001):
        total = 0
        for _ in range(161):
            k = L.orc_randcar_weights(w, 3, C.byref(st))
            assert 0 <= k <= 2
            counts[k] += 1
            total += k + 1
        assert 380 < total < 440
    n = sum(counts)
    for k, p in enumerate((1 / 16, 5 / 16, 10 / 16)):
        assert abs(counts[k] - n * p) < 4.5 * (n * p * (1 - p)) ** 0.5
    assert L.orc_randcar_weights((C.c_uint64 * 2)(0, 0), 2, C.byref(st)) == -1


# ---- infer-* in the oracle's own C restatement (bridge.cl:1126-1281: 7 + 5 + 2 KATs) ----------------------------
def _plist_dict(x):
    """(:HCP (11 15) :C (6 NIL)) as read by sexp -> {'hcp': [11, 15], 'c': [6, None]}"""
    x = unquote(x)
    if not x:
        return {}
    return {str(x[i]).lstrip(":").lower(): x[i + 1] for i in range(0, len(x) - 1, 2)}


def test_infer_kats_in_the_oracle():
    """The oracle's infer-* (used by oracle/ref_build.py, independent of the product's host mirror) replays the
    reference's 14 KATs.  orc_infer_distgen_params computes both halves; each KAT compares the half it is about."""
    full = [[9, 1, 1, 1, 1]] * 4
    n = 0
    for head, keys in (("INFER-HCP-RANGE", ("hcp",)), ("INFER-SUIT-RANGES", ("c", "d", "h", "s"))):
        for where, form, exp in load_kats(head):
            sup = full if form[1] == ["SUPPLY", ["ALL-CARDS"]] else unquote(form[1])
            got = O.infer_distgen_params(sup, _plist_dict(form[2]) or None, [_plist_dict(o) for o in form[3:]])
            want = _plist_dict(exp)
            assert {k: v for k, v in got.items() if k in keys} == want, (where, got, want)
            n += 1
    for where, form, exp in load_kats("INFER-DISTGEN-PARAMS"):
        got = O.infer_distgen_params(unquote(form[1]), _plist_dict(form[2]), [_plist_dict(o) for o in (unquote(form[3]) or [])])
        assert got == _plist_dict(exp) and list(got) == list(_plist_dict(exp)), where
        n += 1
    assert n == 7 + 5 + 2
    with pytest.raises(ValueError):                           # bad-range (bridge.cl:1144-1146)
        O.infer_distgen_params(full, {"hcp": [30, 35]}, [{"hcp": [12, 15]}])
    with pytest.raises(ValueError):                           # bridge.cl:1223-1225
        O.infer_distgen_params(full, {"s": [10, None]}, [{"s": [5, 6]}])


def test_oracle_build_property():
    """bridge.cl:2989-3003 on the oracle's own `build`: three ranged seats + one fixed hand; every hand -- also the
    never-sampled fourth -- has HCP <= 11 and suit lengths 2..5; the four hands tile the deck"""
    rng = {"hcp": [0, 11], "c": [2, 5], "d": [2, 5], "h": [2, 5], "s": [2, 5]}
    fixed = sum(1 << (13 * s + r) for s, ranks in enumerate([[12, 9, 5], [12, 11, 3, 2], [4, 3, 2, 1], [4, 2]]) for r in ranks)
    hands, cnt = O.ref_build_batch([fixed], [rng, rng, rng], 200, seed=3, threads=2)
    assert (cnt == 4).all()
    for row in hands.tolist():
        assert row[0] == fixed and row[0] | row[1] | row[2] | row[3] == (1 << 52) - 1
        for h in row[1:]:
            lens = [bin((h >> (13 * s)) & 0x1FFF).count("1") for s in range(4)]
            hcp = sum(max(r - 8, 0) for s in range(4) for r in range(13) if (h >> (13 * s + r)) & 1)
            assert hcp <= 11 and all(2 <= x <= 5 for x in lens)


def test_trick_source_is_exactly_uniform():
    """orc_tricks_of (mirror of csrc/score.cuh): digits of an accepted word are uniform on 0..13 and independent; the
    accept test is Lemire's (x * 14^8 mod 2^32 >= 2^32 mod 14^8)"""
    import ctypes as C
    import numpy as np
    assert (1 << 32) % 14 ** 8 == 1343389184
    key = (C.c_uint32 * 2)(123, 456)
    out = (C.c_uint32 * 8)()
    counts = np.zeros((8, 14), dtype=np.int64)
    n = 60000
    for i in range(n):
        O.lib().orc_tricks_of(key, i, 7, 0, 1, out)
        for k in range(8):
            counts[k, out[k]] += 1
    exp = n / 14
    assert (abs(counts - exp) < 5 * exp ** 0.5).all()
    # score tallies: rows sum to n; a pair whose difference can exceed 3999 reports the overflow in `dropped`
    cs = [O.Contract(7, 4, 2, 2), O.Contract(7, 4, 0, 2)]          # 7NT redoubled, S and N
    tr, im, dropped = O.score_tally_index(9, 0, 20000, cs, [1, 1], [0, 1])
    assert tr[0].sum() == 20000 and tr[1].sum() == 20000 and im[0].sum() + dropped[0] == 20000 and dropped[0] > 0
