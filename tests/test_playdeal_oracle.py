"""The oracle's restatement of the trick estimator (oracle/playdeal.py) against

* the reference's own 67 inline known-answer tests (bridge.cl:1514-2976; tests/golden/playdeal_kats.json).  Eight of
  them FAIL IN THE REFERENCE ITSELF at this snapshot (stale literals; tests/golden/reference_inline_tests.json is the
  record of the reference's code evaluating its own tests under oracle/minilisp.py): for those the restatement is held
  to the value the reference's code computes;
* outputs of the reference's `play-deal` itself on seeded random deals (tests/golden/playdeal_ref.json, made by
  oracle/run_reference.py);

and the device code of the CUDA path (bridge_mc_b200/csrc/playdeal.cuh) compiled for the host by the test-only
harness tests/playdeal_host_selftest.cpp against both -- the part of the CUDA path's parity that needs no GPU."""
import random
import subprocess
import json
import os
import sys

import numpy as np
import pytest

from bridge_mc_b200 import sexp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import playdeal as P  # noqa: E402
from playdeal_kat_eval import Evaluator, equal  # noqa: E402


def make_instance(cls, **kw):
    cls = str(cls)
    if cls == "HAND":
        return P.Hand(kw["="], kw.get("id"))
    if cls == "TRICK":
        return P.Trick(suit=kw.get("suit"), trump=kw.get("trump"), high=kw.get("high"), winner=kw.get("winner"),
                       remaining=kw.get("remaining"))
    if cls == "CACHE":
        return P.Cache(kw["!"])
    if cls == "PLAY-STRATEGY":
        return P.PlayStrategy(trump=kw.get("trump"), hands=kw.get("hands"))
    if cls == "OUTCOME":
        return P.Outcome(remaining=kw.get("remaining"))
    raise KeyError(cls)


def after_tricks(obj, outcome):
    return obj.after_tricks(outcome) if isinstance(obj, P.PlayStrategy) else P.com_after_tricks(obj, outcome)


def tricks(x):
    return x.tricks


def str2deal(text):
    return [P.str2hand(line) for line in str(text).split("\n")]


FUNCTIONS = {
    "MAKE-INSTANCE": make_instance, "STR2HAND": lambda s: P.str2hand(str(s)), "STR2DEAL": str2deal, "MKHAND": P.mkhand,
    "SUITS": lambda h: h.suits, "HANDID": lambda h: h.id, "TRICKS": tricks, "REMAINING": lambda o: o.remaining,
    "AS-LIST": lambda t: t.as_list(), "AS-SUMMARY": lambda t: t.as_summary(),
    "PLAY-LOW": P.play_low, "CARD-HIGHER": lambda t, a, b: t.card_higher(a, b), "WEAK-SUIT": P.weak_suit,
    "CLOSEST-CARD": P.closest_card, "BEAT-OR-LOW": P.beat_or_low, "FINESSE-IF-POSSIBLE": P.finesse_if_possible,
    "INVERT-IF-NEEDED": P.invert_if_needed, "PLAY-ROUND": P.play_round, "MK-ROUTE": P.mk_route,
    "PLAY-ROUTES": P.play_routes, "DO-NEXT": P.do_next, "SIMPLE-TRICK": P.simple_trick, "AFTER-TRICKS": after_tricks,
    "RUN": lambda cache, trump, a, b, c, d, suits=None: cache.run(trump, a, b, c, d, suits=[P.suitno(s) for s in suits]
                                                                   if suits is not None else [0, 1, 2, 3]),
    "IMMEDIATE-TRICKS": P.immediate_tricks, "FIND-OPP-WEAKNESS": P.find_opp_weakness, "HUNT-CARD": P.hunt_card,
    "PLAY-STRATEGY": P.play_strategy, "PLAY-DEAL": P.play_deal,
}

KATS = json.load(open(os.path.join(HERE, "golden", "playdeal_kats.json"), encoding="utf-8"))
ALL_INLINE = json.load(open(os.path.join(HERE, "golden", "reference_inline_tests.json")))["results"]
INLINE = {r["line"]: r for r in ALL_INLINE if r["file"] == "bridge.cl"}
REF = json.load(open(os.path.join(HERE, "golden", "playdeal_ref.json")))["rows"]
STALE = {2694, 2761, 2768, 2775, 2814, 2868, 2896, 2965}      # fail in the reference itself


def test_fixture_is_complete():
    assert len(KATS) == 67


@pytest.mark.parametrize("kat", KATS, ids=lambda k: f"bridge.cl:{k['line']}")
def test_reference_kat(kat):
    ev = Evaluator(FUNCTIONS)
    got = ev.eval(sexp.read(kat["form"]))
    rec = INLINE[kat["line"]]
    if rec["ok"]:
        want = ev.eval(sexp.read(kat["expected"]))
    else:                       # stale literal: the value the reference's own code computes
        want = sexp.read(rec["reference_computes"])
    if isinstance(got, str) and not isinstance(got, sexp.Sym):
        assert str(got) == str(want)
    else:
        assert equal(got, want), f"bridge.cl:{kat['line']}: {got!r} != {want!r}"


def test_the_reference_fails_exactly_the_stale_tests():
    assert {line for line, r in INLINE.items() if not r["ok"]} == STALE
    assert sum(1 for r in INLINE.values() if r["ok"]) == 159
    # bidding.cl (93) and compscore.cl (15): all green under the reference's own code
    assert sum(1 for r in ALL_INLINE if r["file"] == "bidding.cl" and r["ok"]) == 93
    assert sum(1 for r in ALL_INLINE if r["file"] == "compscore.cl" and r["ok"]) == 15
    assert len(ALL_INLINE) == 275


def _hands(row):
    return [P.Hand(h, id=i) for i, h in enumerate(row["hands"])]


GOOD = [r for r in REF if "tricks" in r]


def test_reference_outputs_fixture():
    assert len(REF) == 240 and len(GOOD) == 238
    assert sum(1 for r in GOOD if sum(len(s) for s in r["hands"][0]) == 13) >= 46      # full deals


@pytest.mark.parametrize("row", GOOD, ids=lambda r: "k%d-t%s" % (sum(len(s) for s in r["hands"][0]), r["trump"]))
def test_oracle_matches_the_reference_run(row):
    o = P.play_deal(row["trump"], *_hands(row))
    assert o.tricks == row["tricks"] and o.score == row["score"] and o.winner_nos == row["winner_nos"]


# --- the CUDA path's device code, compiled for the host ---------------------------------------------------------
@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("pd") / "playdeal_host_selftest")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "playdeal_host_selftest.cpp")], check=True)
    return exe


def run_host(exe, jobs):
    text = "".join("%d %s\n" % (-1 if t is None else t, " ".join(str(sum(1 << r for r in s)) for h in hs for s in h)) for hs, t in jobs)
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.strip().split("\n")
    res = []
    for line in out:
        f = line.split()
        n = int(f[3])
        cards = [int(x, 16) for x in f[6:6 + 4 * n]]
        res.append({"status": int(f[0]), "score": [int(f[1]), int(f[2])],
                    "tricks": [[None if c == 255 else [c >> 4, c & 15] for c in cards[4 * t:4 * t + 4]] for t in range(n)],
                    "winner_nos": [int(x) for x in f[6 + 4 * n:]], "arena_peak": int(f[4])})
    return res


def test_device_code_on_host_matches_the_reference_run(host_binary):
    res = run_host(host_binary, [(r["hands"], r["trump"]) for r in GOOD])
    for r, g in zip(GOOD, res):
        assert g["status"] == 0
        assert g["tricks"] == r["tricks"] and g["score"] == r["score"] and g["winner_nos"] == r["winner_nos"]


def _rand_deal(rng, k):
    cards = rng.sample(range(52), 4 * k)
    return [[[c % 13 for c in sorted(cards[h * k:(h + 1) * k]) if c // 13 == s] for s in range(4)] for h in range(4)]


def test_device_code_on_host_matches_the_oracle_on_random_deals(host_binary):
    rng = random.Random(7)
    jobs = [(_rand_deal(rng, k), rng.choice([None, 0, 1, 2, 3])) for k in (1, 2, 3, 4, 5, 6, 7, 9, 11) for _ in range(24)]
    jobs += [(_rand_deal(rng, 13), rng.choice([None, 0, 1, 2, 3])) for _ in range(12)]
    res = run_host(host_binary, jobs)
    for (hs, t), g in zip(jobs, res):
        o = P.play_deal(t, *[P.Hand(h, id=i) for i, h in enumerate(hs)])
        assert g["status"] == 0 and g["tricks"] == o.tricks and g["score"] == o.score and g["winner_nos"] == o.winner_nos


def test_device_code_under_address_and_ub_sanitizers(tmp_path):
    """compute-sanitizer is closed on the GPU pool; the estimator's memory discipline (arena offsets, the shared
    temporaries' stack, the fixed-size route / guard / target tables and their overflow paths) is the same code on the
    host, so it is run under ASan + UBSan here: normal arenas, and arenas small enough that many deals overflow (status 1
    must be reported, nothing may be touched out of bounds)."""
    exe = str(tmp_path / "pd_asan")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe,
                    os.path.join(HERE, "playdeal_host_selftest.cpp")], check=True)
    rng = random.Random(23)
    jobs = [(_rand_deal(rng, 13), rng.choice([None, 0, 1, 2, 3])) for _ in range(40)] + [(_rand_deal(rng, k), None) for k in (1, 2, 5, 9)]
    text = "".join("%d %s\n" % (-1 if t is None else t, " ".join(str(sum(1 << r for r in s)) for h in hs for s in h)) for hs, t in jobs)
    full = subprocess.run([exe, str(8 << 20)], input=text, capture_output=True, text=True)
    assert full.returncode == 0, full.stderr[-2000:]
    assert all(line.split()[0] == "0" for line in full.stdout.strip().split("\n"))
    small = subprocess.run([exe, str(48 << 10)], input=text, capture_output=True, text=True)
    assert small.returncode == 0, small.stderr[-2000:]
    status = [line.split()[0] for line in small.stdout.strip().split("\n")]
    assert set(status) <= {"0", "1"} and status.count("1") >= 5
    for a, b in zip(full.stdout.strip().split("\n"), small.stdout.strip().split("\n")):
        if b.split()[0] == "0":
            assert a.split()[:4] == b.split()[:4] and a.split()[6:] == b.split()[6:]      # same outcome when the arena suffices


def test_scheduling_probe_ranks_deals_by_measured_length(host_binary):
    """play_deal's probe mode (the launcher plays the deals it scores highest first) against the device cycles measured
    for 32 559 deals on a B200 (profiles/r02_playdeal_order_train.npz; the model was fitted on them by
    tools/fit_playdeal_order.py): the score must rank the deals well, and most of the hundred longest must sit in its top decile.
    The order never changes a result (tests/test_gpu_playdeal.py); this guards the speed-up against a stale model."""
    from scipy.stats import spearmanr
    d = np.load(os.path.join(HERE, "..", "profiles", "r02_playdeal_order_train.npz"))
    hands, cyc = d["hands"][:6000], d["cycles"][:6000]
    text = "".join("-1 " + " ".join(str((int(h) >> (16 * s)) & 0x1FFF) for h in hs for s in range(4)) + "\n" for hs in hands)
    out = subprocess.run([host_binary, str(64 << 20), "probe"], input=text, capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert len(out) == len(hands) and all(line.split()[0] == "0" and len(line.split()) == 2 + 74 for line in out)
    score = np.array([int(line.split()[1]) for line in out], dtype=np.float64)
    assert spearmanr(score, cyc)[0] > 0.7
    top = set(np.argsort(-score)[:600].tolist())
    assert sum(int(i) in top for i in np.argsort(-cyc)[:100]) >= 50


FROZEN_DIGEST = "6da7e2647984f06fe253df57709895e5ad9811681f2f9cb1d75f1fcb6f023f05"


def frozen_deals():
    """1 000 seeded full deals and 500 endings of 3 / 5 / 8 / 10 cards, each with a random trump (or none): the set whose
    outcomes are frozen in FROZEN_DIGEST (also used by tests/test_gpu_playdeal.py)."""
    rng, jobs = random.Random(77), []
    for i in range(1500):
        d = list(range(52))
        rng.shuffle(d)
        k = 13 if i < 1000 else rng.choice([3, 5, 8, 10])
        masks = [sum(1 << (c % 13) for c in d[k * j:k * j + k] if c // 13 == s) for j in range(4) for s in range(4)]
        jobs.append((rng.choice([0, 1, 2, 3, -1]), masks))
    return jobs


def frozen_digest(rows):
    """rows: (status, score0, score1, n, cards[n][4], winners[n])"""
    import hashlib
    text = "\n".join(" ".join([str(st), str(s0), str(s1), str(n)] + ["%02x" % c for t in cards for c in t] + [str(w) for w in wno])
                     for st, s0, s1, n, cards, wno in rows)
    return hashlib.sha256(text.encode()).hexdigest()


def test_outcomes_of_the_frozen_deal_set(host_binary):
    """Every card of 1 500 deals and endings with random trumps, as a digest.  The digest was taken from the build that the
    tests above tie to the reference's own run and to the restatement, before the estimator was restructured (the trick in
    registers, memoised look-aheads, route masks, the position cache), and every later build has to reproduce it."""
    text = "".join(f"{t} " + " ".join(map(str, m)) + "\n" for t, m in frozen_deals())
    out = subprocess.run([host_binary], input=text, capture_output=True, text=True, check=True).stdout.strip().split("\n")
    rows = []
    for line in out:
        f = line.split()
        n = int(f[3])
        assert f[0] == "0"
        cards = [[int(x, 16) for x in f[6 + 4 * t:10 + 4 * t]] for t in range(n)]
        rows.append((0, int(f[1]), int(f[2]), n, cards, [int(x) for x in f[6 + 4 * n:6 + 5 * n]]))
    assert frozen_digest(rows) == FROZEN_DIGEST
