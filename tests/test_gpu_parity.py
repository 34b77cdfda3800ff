"""GPU parity tests: every call goes through the C ABI (libbridgemc.so) and is
checked against the CPU oracle (oracle/bmc_oracle.c) on the same inputs."""
import numpy as np
import pytest

from bridge_mc_b200 import bidding
from bridge_mc_b200.cards import seat_masks
from bridge_mc_b200.engine import Constraints, seat_spec
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x6272696467656D63


def sort_records(rec):
    rec = np.ascontiguousarray(rec, dtype=np.uint64).reshape(-1, 2)
    order = np.lexsort((rec[:, 0], rec[:, 1]))
    return rec[order]


def oracle_features(rec):
    """int32[n,4,11] assess-hand vectors of the four seats"""
    m = seat_masks(rec)
    return O.features_batch(m.reshape(-1)).reshape(-1, 4, 11)


@pytest.mark.parametrize("first,n", [(0, 1), (0, 31), (5, 1000), (123456789012, 200_000)])
def test_sampler_bit_exact_vs_cpu_mirror(engine, first, n):
    rec, tal, st = engine.sample(None, n_raw=n, seed=SEED, first_index=first)
    ref, void = O.gpu_deal_batch(SEED, first, n)
    assert tal["raw"] == n
    assert tal["voided"] == int(void.sum())
    assert tal["accepted"] == n - int(void.sum()) == rec.shape[0]
    assert np.array_equal(sort_records(rec), sort_records(ref[void == 0]))


def test_every_record_is_a_deal(engine):
    rec, _, _ = engine.sample(None, n_raw=100_000, seed=7)
    m = seat_masks(rec)
    pc = np.zeros(m.shape, dtype=np.int64)
    for b in range(64):
        pc += ((m >> np.uint64(b)) & np.uint64(1)).astype(np.int64)
    assert (pc == 13).all()
    assert ((m[:, 0] | m[:, 1] | m[:, 2] | m[:, 3]) == np.uint64(0x1FFF1FFF1FFF1FFF)).all()


def test_features_bit_exact(engine):
    rec, _, _ = engine.sample(None, n_raw=100_000, seed=11)
    acc, feats = engine.filter(rec)
    ref = oracle_features(rec)
    assert (acc == 1).all()
    assert np.array_equal(feats["len"].astype(np.int32), ref[:, :, 0:4])
    assert np.array_equal(feats["suit_hcp"].astype(np.int32), ref[:, :, 4:8])
    assert np.array_equal(feats["hcp"].astype(np.int32), ref[:, :, 8])
    assert np.array_equal(feats["losers"].astype(np.int32), ref[:, :, 9])
    assert np.array_equal(feats["balanced"].astype(np.int32), ref[:, :, 10])


@pytest.mark.parametrize("scheme_id,table", [(0, bidding.openings), (1, bidding.nt_responses(1)),
                                             (2, bidding.nt_responses(2)), (3, bidding.two_c_responses)])
def test_choose_bid_bit_exact(engine, scheme_id, table):
    rec, _, _ = engine.sample(None, n_raw=60_000, seed=13 + scheme_id)
    _, feats = engine.filter(rec, None, table.scheme())
    ref = O.choose_bid_batch(scheme_id, seat_masks(rec).reshape(-1)).reshape(-1, 4)
    assert np.array_equal(feats["bid"].astype(np.int32), ref)
    assert len(np.unique(ref)) > 3


def test_accept_flags_1nt_south(engine):
    rule = bidding.openings.form((1, "NT"))
    cons = Constraints(None, [None, None, rule, None])
    rec, _, _ = engine.sample(None, n_raw=200_000, seed=17)
    acc, _ = engine.filter(rec, cons, None, want_features=False)
    south = seat_masks(rec)[:, 2]
    f = O.features_batch(south)
    ref = np.array([O.lib().orc_test_rule(0, 2, (O.C.c_int * 11)(*row)) for row in f.tolist()], dtype=np.uint8)
    assert np.array_equal(acc, ref)
    assert 0.03 < acc.mean() < 0.046


def mirror_stream(cons, seed, first, n):
    """CPU mirror of the raw stream the sampler draws for `cons` (after its first sampling call):
    (records, void_a, void_b)"""
    if cons is not None and cons.mode == 2:
        rec, fl = O.gpu_deal_seatfirst_batch(seed, first, n, cons.primary, cons.primary_hcp)
        return rec, (fl & 3) != 0, (fl & 4) != 0
    rec, void = O.gpu_deal_batch(seed, first, n)
    return rec, void.astype(bool), np.zeros(n, dtype=bool)


@pytest.mark.parametrize("mode", [1, 2])
def test_constrained_sampler_equals_filtered_stream(engine, mode):
    """rejection sampling = filtering the raw stream: same accepted set, in both dealing modes"""
    rule = bidding.openings.form((1, "NT"))
    cons = Constraints(None, [None, None, rule, None]).set_mode(mode)
    n = 300_000
    got, tal, _ = engine.sample(cons, n_raw=n, seed=SEED, first_index=1000)
    assert cons.mode == mode and cons.primary == 2
    ref, va, vb = mirror_stream(cons, SEED, 1000, n)
    acc, _ = engine.filter(ref, cons, None, want_features=False)
    keep = (acc == 1) & ~va & ~vb
    assert tal["raw"] == n
    assert tal["accepted"] == int(keep.sum()) == got.shape[0]
    assert np.array_equal(sort_records(got), sort_records(ref[keep]))
    # voided = indices discarded by a Lemire check in a stage that was actually executed
    if mode == 1:
        assert tal["voided"] == int(va.sum())
    else:
        assert cons.primary_hcp == (15, 17)
        _, fl = O.gpu_deal_seatfirst_batch(SEED, 1000, n, 2, (15, 17))
        f = O.features_batch(seat_masks(ref)[:, 2])
        v1, v2, v3 = (fl & 1) != 0, (fl & 2) != 0, (fl & 4) != 0
        pass_a1 = (f[:, 8] >= 15) & (f[:, 8] <= 17) & ~v1            # stage A1 tests the HCP window only
        assert np.array_equal(pass_a1, (fl & 8) != 0)
        expect = v1.sum() + (~v1 & pass_a1 & v2).sum() + (~v1 & ~v2 & (acc == 1) & v3).sum()
        assert tal["voided"] == int(expect)      # this rule is exactly its ranges: passing stage A2 == accepted


def test_seat_first_void_count_of_stage_a1(engine):
    """stage A1 voids u >= S * C(52,13) (S = floor(2^64 / C(52,13))): 1.22e-8 per index.  With a window no
    hand reaches before stage A2 matters (37 HCP: 4 hands in 6.35e11) every void counted is an A1 void."""
    cons = Constraints([None, None, seat_spec(hcp=(37, 37)), None]).set_mode(2)
    n = 1 << 34
    _, tal, _ = engine.sample(cons, n_raw=n, seed=77, cap=0)
    n_hands = 635013559600
    p = 1 - (2 ** 64 // n_hands) * n_hands / 2 ** 64
    assert tal["raw"] == n and tal["accepted"] <= 2
    assert abs(tal["voided"] - n * p) < 5 * (n * p) ** 0.5, (tal["voided"], n * p)


@pytest.mark.parametrize("seat,spec", [
    (0, dict(hcp=(0, 37))),                        # every class admitted: theta = S * C(52,13)
    (1, dict(hcp=(20, 21))),
    (3, dict(hcp=(0, 7), s=(5, 13))),
    (2, dict(hcp=(11, 14), h=(4, 6), d=(0, 3))),
    (0, dict(hcp=(28, 37))),                       # a handful of tiny classes
    (1, dict(c=(6, 13))),                          # no HCP range at all
])
def test_seat_first_windows_and_seats_vs_mirror(engine, seat, spec):
    """rank draw + unranking for other HCP windows and primary seats: the accepted set equals the CPU
    mirror's stream filtered by the same constraint"""
    specs = [None] * 4
    specs[seat] = seat_spec(**spec)
    cons = Constraints(specs).set_mode(2)
    n = 150_000
    first = 3 if seat % 2 == 0 else (1 << 45) + 7                                    # odd first index: a split pair
    got, tal, _ = engine.sample(cons, n_raw=n, seed=SEED + seat, first_index=first)
    assert (cons.mode, cons.primary) == (2, seat)
    lo, hi = cons.primary_hcp
    assert (lo, hi) == (spec.get("hcp", (0, 37))[0], min(spec.get("hcp", (0, 37))[1], 37))
    ref, va, vb = mirror_stream(cons, SEED + seat, first, n)
    acc, _ = engine.filter(ref, cons, None, want_features=False)
    keep = (acc == 1) & ~va & ~vb
    assert tal["raw"] == n and tal["accepted"] == int(keep.sum()) == got.shape[0]
    assert np.array_equal(sort_records(got), sort_records(ref[keep]))


def test_chi_square_seat_first_full_distribution(engine):
    """the whole rank / unrank path at acceptance 1: with an all-admitting window on South the seat-first
    kernel must reproduce the exact HCP and sorted-shape pmfs of every seat (parity rejected if p < 1e-3)"""
    from math import comb
    from itertools import permutations
    from bridge_mc_b200._lib import shape_table
    cons = Constraints([None, None, seat_spec(hcp=(0, 37)), None]).set_mode(2)
    n = 20_000_000
    _, tal, _ = engine.sample(cons, n_raw=n, seed=606, cap=0)
    acc = tal["accepted"]
    assert cons.mode == 2 and 0.99 * n < acc <= n
    hcp_num, _ = O.Distgen(c=(0, 13), d=(0, 13), h=(0, 13), s=(0, 13)).pmf()
    total = comb(52, 13)
    exp_hcp = np.array(hcp_num, float) / total * acc
    shape_w = [sum(comb(13, p[0]) * comb(13, p[1]) * comb(13, p[2]) * comb(13, p[3]) for p in set(permutations(r)))
               for r in shape_table().tolist()]
    exp_shape = np.array(shape_w, float) / total * acc
    exp_len = np.array([comb(13, l) * comb(39, 13 - l) for l in range(14)], float) / total * acc
    for k in range(4):
        assert _chi2_p(tal["hcp"][k], exp_hcp) > 1e-3
        assert _chi2_p(tal["shape"][k], exp_shape) > 1e-3
        for su in range(4):
            assert _chi2_p(tal["len"][k][su], exp_len) > 1e-3


def test_auto_mode_picks_by_pass_rate(engine):
    tight = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    loose = Constraints([None, seat_spec(hcp=(0, 20)), None, None])
    engine.sample(tight, n_raw=1000, cap=0)
    engine.sample(loose, n_raw=1000, cap=0)
    assert (tight.mode, tight.primary) == (4, 2)             # exact acceptance 0.0382 < 0.30: seat-first with the class table
    assert (loose.mode, loose.primary) == (1, 1)
    assert tight.primary_acceptance[0] == 24235203726 and loose.primary_acceptance[0] > 0.9 * 635013559600
    # a shapeless, wide window: still under 0.30, a few hundred thousand classes
    wide = Constraints([None, None, seat_spec(hcp=(0, 7)), None])
    engine.sample(wide, n_raw=1000, cap=0)
    hands, classes = wide.primary_acceptance
    assert wide.mode == 4 and hands / 635013559600 < 0.30 and classes > 100_000


def test_seat_first_multi_seat_and_programs(engine):
    """seat-first with rule programs on every seat (C3-style), against the CPU mirror + filter"""
    p = bidding.openings.passes()
    cons = Constraints(None, [p, p, bidding.openings.chooses((1, "NT")), p]).set_mode(2)
    n = 400_000
    got, tal, _ = engine.sample(cons, n_raw=n, seed=21)
    ref, va, vb = mirror_stream(cons, 21, 0, n)
    acc, _ = engine.filter(ref, cons, None, want_features=False)
    keep = (acc == 1) & ~va & ~vb
    assert np.array_equal(sort_records(got), sort_records(ref[keep]))
    assert 0.012 < keep.mean() < 0.020                      # P-P-1NT-P: 0.01605


def test_tallies_match_features(engine):
    cons = Constraints([None, None, seat_spec(hcp=(12, 20), s=(5, None)), None])
    rec, tal, _ = engine.sample(cons, n_raw=400_000, seed=23)
    _, feats = engine.filter(rec)
    assert tal["accepted"] == rec.shape[0] > 1000
    for k in range(4):
        assert np.array_equal(tal["hcp"][k], np.bincount(feats["hcp"][:, k], minlength=38)[:38])
        for s in range(4):
            assert np.array_equal(tal["len"][k, s], np.bincount(feats["len"][:, k, s], minlength=14))
    assert (feats["hcp"][:, 2] >= 12).all() and (feats["hcp"][:, 2] <= 20).all() and (feats["len"][:, 2, 3] >= 5).all()
    from bridge_mc_b200._lib import shape_table
    shapes = [tuple(r) for r in shape_table().tolist()]
    for k in range(4):
        srt = np.sort(feats["len"][:, k, :], axis=1)[:, ::-1]
        cls = np.array([shapes.index(tuple(r)) for r in srt.tolist()])
        assert np.array_equal(tal["shape"][k], np.bincount(cls, minlength=39))


def test_scores_bit_exact_full_table(engine):
    from bridge_mc_b200._lib import Contract
    L = O.lib()
    contracts, vul, meta = [], [], []
    for suit in range(5):
        for level in range(0, 8):
            for dbl in range(3):
                for decl in (-1, 0, 1, 2, 3):
                    for v in (0, 1):
                        contracts.append(Contract(level, suit, decl, dbl))
                        vul.append(v)
                        meta.append((level, suit, decl, dbl, v))
    for lo in range(0, len(contracts), 8):
        cs, vs, ms = contracts[lo:lo + 8], vul[lo:lo + 8], meta[lo:lo + 8]
        tricks = np.repeat(np.arange(14, dtype=np.uint8)[:, None], len(cs), axis=1)
        got = engine.score(cs, vs, tricks)
        for j, (level, suit, decl, dbl, v) in enumerate(ms):
            for t in range(14):
                assert got[t, j] == L.orc_base_score(level, suit, decl, dbl, t, v), (level, suit, decl, dbl, v, t)


def test_match_points_bit_exact(engine):
    rng = np.random.default_rng(5)
    a = rng.integers(-7600, 2981, size=50_000).astype(np.int32)
    b = rng.integers(-7600, 2981, size=50_000).astype(np.int32)
    imps, status = engine.match_points(a, b)
    for i in range(0, 50_000, 7):
        ref = O.imps(int(a[i]) - int(b[i]))
        if ref is None:
            assert status[i] == 1
        else:
            assert status[i] == 0 and imps[i] == ref


def _chi2_p(obs, exp):
    from scipy.stats import chi2
    obs, exp = np.asarray(obs, float), np.asarray(exp, float)
    # merge bins until every expected count >= 5
    order = np.argsort(exp)
    o, e, mo, me = [], [], 0.0, 0.0
    for i in order:
        mo += obs[i]; me += exp[i]
        if me >= 5:
            o.append(mo); e.append(me); mo = me = 0.0
    if me > 0 and e:
        o[-1] += mo; e[-1] += me
    o, e = np.array(o), np.array(e)
    stat = ((o - e) ** 2 / e).sum()
    return chi2.sf(stat, len(o) - 1)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_chi_square_unconstrained(engine, seed):
    """HCP (38 bins) and sorted shape (39 bins) of every seat against the exact pmfs;
    parity is rejected if p < 1e-3."""
    from math import comb
    from bridge_mc_b200._lib import shape_table
    n = 20_000_000
    _, tal, _ = engine.sample(None, n_raw=n, seed=seed, cap=0)
    acc = tal["accepted"]
    hcp_num, _ = O.Distgen(c=(0, 13), d=(0, 13), h=(0, 13), s=(0, 13)).pmf()
    total = comb(52, 13)
    assert sum(hcp_num) == total
    exp_hcp = np.array(hcp_num, float) / total * acc
    shape_w = []
    for a, b, c, d in shape_table().tolist():
        from itertools import permutations
        shape_w.append(sum(comb(13, p[0]) * comb(13, p[1]) * comb(13, p[2]) * comb(13, p[3])
                           for p in set(permutations((a, b, c, d)))))
    assert sum(shape_w) == total
    exp_shape = np.array(shape_w, float) / total * acc
    for k in range(4):
        assert _chi2_p(tal["hcp"][k], exp_hcp) > 1e-3
        assert _chi2_p(tal["shape"][k], exp_shape) > 1e-3


def test_c2_acceptance_rate_exact(engine):
    """South `test-bid (1 NT)`: acceptance 24 235 203 726 / C(52,13) and the exact
    15/16/17 split, within 4 sigma."""
    rule = bidding.openings.form((1, "NT"))
    cons = Constraints(None, [None, None, rule, None])
    n = 50_000_000
    cons.set_mode(1)                      # full deal: every void index is seen, so raw - voided is exact
    _, tal, _ = engine.sample(cons, n_raw=n, seed=99, cap=0)
    valid = tal["raw"] - tal["voided"]
    p = 24235203726 / 635013559600
    sigma = (valid * p * (1 - p)) ** 0.5
    assert abs(tal["accepted"] - valid * p) < 4 * sigma
    split = np.array([9820068156, 8126568540, 6288567030], float)
    exp = split / split.sum() * tal["accepted"]
    obs = tal["hcp"][2][15:18]
    assert obs.sum() == tal["accepted"]
    assert _chi2_p(obs, exp) > 1e-3


def test_c2_acceptance_rate_exact_seat_first(engine):
    """same exact acceptance in seat-first mode; the survival factors of the Lemire checks are
    computed exactly from the digit schedules"""
    from math import prod
    rule = bidding.openings.form((1, "NT"))
    cons = Constraints(None, [None, None, rule, None]).set_mode(2)
    n = 50_000_000
    _, tal, _ = engine.sample(cons, n_raw=n, seed=123, cap=0)
    n_hands = 635013559600
    surv_a = (2 ** 64 // n_hands) * n_hands / 2 ** 64                 # stage A1: u beyond S * C(52,13); A2 draws nothing
    surv_b = prod(1 - ((1 << 32) % prod(m for m in range(39 - 4 * w - 3, 39 - 4 * w + 1) if m >= 2)) / 2 ** 32
                  for w in range(10))
    p = 24235203726 / 635013559600 * surv_a * surv_b
    sigma = (n * p * (1 - p)) ** 0.5
    assert abs(tal["accepted"] - n * p) < 4 * sigma
    split = np.array([9820068156, 8126568540, 6288567030], float)
    obs = tal["hcp"][2][15:18]
    assert obs.sum() == tal["accepted"]
    assert _chi2_p(obs, split / split.sum() * tal["accepted"]) > 1e-3


@pytest.mark.parametrize("seed", [4, 5])
def test_chi_square_seat_first_other_seats(engine, seed):
    """seat-first deals: given South's 1NT hand the other three hands must be uniform -- their HCP
    sums to 40 - South's and their suit lengths follow the exact hypergeometric marginals; here:
    the three seats are exchangeable (pairwise chi-square homogeneity of HCP and shape tallies)."""
    rule = bidding.openings.form((1, "NT"))
    cons = Constraints(None, [None, None, rule, None]).set_mode(2)
    _, tal, _ = engine.sample(cons, n_raw=100_000_000, seed=seed, cap=0)
    from scipy.stats import chi2_contingency
    for key in ("hcp", "shape"):
        table = np.stack([tal[key][0], tal[key][1], tal[key][3]])
        table = table[:, table.sum(axis=0) >= 15]
        assert chi2_contingency(table)[1] > 1e-3
