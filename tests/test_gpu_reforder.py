"""Reference-order mode (bmc_constraints_set_reference_order): the GPU must reproduce the DISTRIBUTION of the
reference's own sequential `build` (bridge.cl:1292-1319), which differs from the uniform distribution over valid
deals once two or more seats are ranged.  Checked against (a) exact pmfs where one seat is ranged and (b) the
reference algorithm restated in the oracle (oracle/ref_build.py) for multi-seat defs."""
import numpy as np
import pytest
from scipy.stats import chi2

from bridge_mc_b200.cards import Hand, seat_masks, str2hand
from bridge_mc_b200.engine import Constraints, seat_spec
from oracle import oracle as O
from oracle.ref_build import ref_build_batch

pytestmark = pytest.mark.gpu


def chi2_p(obs, exp):
    obs, exp = np.asarray(obs, float), np.asarray(exp, float)
    order = np.argsort(exp)
    o, e, mo, me = [], [], 0.0, 0.0
    for i in order:
        mo += obs[i]; me += exp[i]
        if me >= 5:
            o.append(mo); e.append(me); mo = me = 0.0
    if me > 0 and e:
        o[-1] += mo; e[-1] += me
    if len(o) < 2:
        return 1.0
    o, e = np.array(o), np.array(e)
    return chi2.sf(((o - e) ** 2 / e).sum(), len(o) - 1)


def lanes_from52(m52):
    return int(O.mask52_to_lanes64(np.array([m52], dtype=np.uint64))[0])


def test_single_def_matches_exact_pmf(engine):
    spec = dict(hcp=(15, 17), c=(2, 5), d=(2, 5), h=(2, 4), s=(2, 4))
    cons = Constraints([seat_spec(**spec)]).set_reference_order(1)
    n = 3_000_000
    _, tal, _ = engine.sample(cons, n_raw=n, seed=5, cap=0)
    assert cons.mode == 3 and tal["raw"] == n == tal["accepted"] and tal["voided"] == 0      # every index yields a deal
    hcp_num, len_num = O.Distgen(**spec).pmf()
    tot = float(sum(hcp_num))
    assert chi2_p(tal["hcp"][0], np.array(hcp_num[:38]) / tot * n) > 1e-3
    for s in range(4):
        assert chi2_p(tal["len"][0][s], np.array(len_num[s]) / tot * n) > 1e-3
    assert tal["hcp"][0][:15].sum() == 0 and tal["hcp"][0][18:].sum() == 0


def test_empty_def_has_the_low_card_cap(engine):
    """(make-instance 'deal) / a def of NIL: the sampled seat can never hold a suit longer than the nine low cards
    (bridge.cl:1000-1001); the deal-all seats can."""
    cons = Constraints([seat_spec()]).set_reference_order(1)
    n = 40_000_000
    _, tal, _ = engine.sample(cons, n_raw=n, seed=6, cap=0)
    assert tal["accepted"] == n
    assert tal["len"][0][:, 10:].sum() == 0
    assert tal["len"][1][:, 10:].sum() + tal["len"][2][:, 10:].sum() + tal["len"][3][:, 10:].sum() > 0
    hcp_num, len_num = O.Distgen().pmf()                      # reference nil bounds: 635 002 871 360 hands
    tot = float(sum(hcp_num))
    assert int(tot) == 635002871360
    assert chi2_p(tal["hcp"][0], np.array(hcp_num[:38]) / tot * n) > 1e-3
    for s in range(4):
        assert chi2_p(tal["len"][0][s], np.array(len_num[s]) / tot * n) > 1e-3


def test_lone_def_drops_its_suit_hcp_range(engine):
    """bridge.cl:1227-1229: without later defs the third element of a suit spec is spliced away"""
    cons = Constraints([seat_spec(s=(5, None, (8, None)))]).set_reference_order(1)
    rec, tal, _ = engine.sample(cons, n_raw=20000, seed=7)
    _, feats = engine.filter(rec)
    assert (feats["len"][:, 0, 3] >= 5).all()
    assert (feats["suit_hcp"][:, 0, 3] < 8).any()


N_ORACLE = 100_000          # builds of the reference algorithm per comparison (C restatement, all host threads)


def _oracle_sample(consts, defs, n, seed=1):
    """n builds of the reference's own algorithm (oracle/ref_build.py: distgen + infer-* + build, all C); the builds
    for which the reference signals an error are left out, as the GPU counts them in `voided`"""
    hands, n_err = ref_build_batch(consts, defs, n, seed)
    return O.mask52_to_lanes64(hands.reshape(-1)).reshape(-1, 4), n_err


def _compare(engine, consts_hands, defs, n_oracle, seats_sampled):
    specs = [seat_spec(fixed=h) for h in consts_hands] + [seat_spec(**d) for d in defs]
    cons = Constraints(specs).set_reference_order(len(defs))
    n = 4_000_000
    _, tal, _ = engine.sample(cons, n_raw=n, seed=9, cap=0)
    assert tal["accepted"] + tal["voided"] == n and tal["accepted"] > 0.9 * n
    consts52 = [sum(1 << c for c in h.cards()) for h in consts_hands]
    ora, n_err = _oracle_sample(consts52, defs, n_oracle)
    # the share of builds that signal an error agrees too (binomial, 5 sigma)
    p_void = tal["voided"] / n
    assert abs(n_err - p_void * n_oracle) <= 5 * (max(p_void, 1 / n_oracle) * n_oracle) ** 0.5 + 1, (n_err, p_void)
    f = O.features_batch(ora.reshape(-1)).reshape(ora.shape[0], -1, 11)
    ps = []
    for k in range(4):
        gpu_pmf = tal["hcp"][k] / tal["accepted"]
        obs = np.bincount(f[:, k, 8], minlength=38)[:38]
        ps.append(chi2_p(obs, gpu_pmf * obs.sum()))
        for s in range(4):
            gpu_len = tal["len"][k][s] / tal["accepted"]
            ps.append(chi2_p(np.bincount(f[:, k, s], minlength=14), gpu_len * ora.shape[0]))
    assert min(ps) > 1e-4, ps            # 4 x (1 + 4) tests
    return tal, f


def test_two_ranged_seats_follow_the_sequential_distribution(engine):
    defs = [dict(hcp=(15, 17), s=(5, None)), dict(hcp=(10, 11), h=(4, None))]
    tal, f = _compare(engine, [], defs, N_ORACLE, 2)
    # and this is NOT the uniform distribution over valid deals: joint rejection weights seat 0's hands by
    # the number of completions, which shifts its tallies
    cons_joint = Constraints([seat_spec(hcp=(15, 17), s=(5, 13)), seat_spec(hcp=(10, 11), h=(4, 13))])
    _, tj, _ = engine.sample(cons_joint, n_raw=400_000_000, seed=9, cap=0)
    pj = tj["len"][0][2] / tj["accepted"]                              # seat 0's heart length
    ps = tal["len"][0][2] / tal["accepted"]
    assert np.abs(pj - ps).max() > 0.01


def test_fixed_hand_and_three_ranged_seats(engine):
    """the reference's own property-test shape (bridge.cl:2989-3003): lower bounds are inferred because the four
    seats are all defined, and the last ranged seat is implied, never sampled"""
    rng = dict(hcp=(0, 11), c=(2, 5), d=(2, 5), h=(2, 5), s=(2, 5))
    fixed = Hand([[12, 9, 5], [12, 11, 3, 2], [4, 3, 2, 1], [4, 2]])
    tal, f = _compare(engine, [fixed], [rng, rng, rng], N_ORACLE, 2)
    for k in (1, 2, 3):                                                # incl. the never-sampled last seat
        assert tal["hcp"][k][12:].sum() == 0
        assert tal["len"][k][:, :2].sum() == 0 and tal["len"][k][:, 6:].sum() == 0


def test_reference_order_errors(engine):
    from bridge_mc_b200._lib import BridgeMcError
    with pytest.raises(BridgeMcError):
        Constraints(None, ["(> hcp 10)"]).set_reference_order(1)             # ranges only
    h = str2hand("♣ AJ7 ♦ AK54 ♥ 6543 ♠ 64")
    with pytest.raises(BridgeMcError):
        Constraints([None, seat_spec(fixed=h)]).set_reference_order(1)        # fixed hands come first
    from bridge_mc_b200._lib import BadRange
    with pytest.raises(BadRange):
        Constraints([seat_spec(hcp=(20, 25)), seat_spec(hcp=(25, 30))])       # impossible for every deal: at creation
    # narrowing fails only for SOME first hands (seat 0 holds 21+ -> fewer than 19 HCP remain): the reference
    # signals bad-range for those builds; here they are counted in `voided`
    cons = Constraints([seat_spec(hcp=(20, 22)), seat_spec(hcp=(19, None))]).set_reference_order(2)
    rec, tal, _ = engine.sample(cons, n_raw=200_000, seed=3)
    assert tal["accepted"] > 0 and tal["voided"] > 0 and tal["accepted"] + tal["voided"] == 200_000
    _, feats = engine.filter(rec)
    assert (feats["hcp"][:, 0] >= 20).all() and (feats["hcp"][:, 0] <= 22).all() and (feats["hcp"][:, 1] >= 19).all()
    assert (feats["hcp"][:, 0] + feats["hcp"][:, 1] <= 40).all()


def test_three_ranged_seats_with_suit_hcp_ranges(engine):
    """three defs with per-suit HCP ranges (third elements), open ends and a seat that lacks keys the others have:
    exercises the suit-HCP narrowing against the later defs and the `others missing` cases of the folds"""
    defs = [dict(hcp=(11, 14), s=(5, None, (4, None)), h=(0, 3)),
            dict(hcp=(6, 9), s=(0, 3, (1, None)), d=(4, None)),
            dict(c=(3, 6))]
    tal, f = _compare(engine, [], defs, N_ORACLE, 3)
    assert tal["hcp"][0][:11].sum() == 0 and tal["hcp"][0][15:].sum() == 0
    assert tal["len"][2][0][:3].sum() == 0 and tal["len"][2][0][7:].sum() == 0      # third seat: 3..6 clubs
