"""Host-side mirror of the reference's trick-estimator entry points, backed by the GPU (libbridgemc.so).

    play-deal        bridge.cl:2961-2963    play_deal(trump, declarer, left, dummy, right) -> Outcome
    best-tricks      bridge.cl:3027-3028    best_tricks(trump, n, e, s, w) -> int
    trump-length     bridge.cl:3083-3086    trump_length(trump, n, e, s, w)
    longest-suit     bridge.cl:3030-3036    longest_suit(*hands)
    best-tricks*     bridge.cl:3038-3043    best_tricks_star(_, n, e, s, w)   (see the note below)
    best-outcomes    bridge.cl:3071-3081    best_outcomes(_, n, e, s, w)

Same names, argument meaning and results as the reference: `trump` is a suit symbol ('c 'd 'h 's / "C".."S"), nil /
'nt for no trump; hands are `Hand`s (four ascending rank lists) and may be partial (endings); the outcome carries
`tricks` (cards in play order as (suit rank)), `winner_nos` and `score` = (tricks of the side on lead first, tricks of
the declaring side).  One call = one GPU launch; the batched forms (`Engine.play_deal`, `Engine.best_tricks`) are what
a sampling job uses.

`best-tricks*` in the reference calls `(longest-suit (list n e s w))` on a &rest function, which signals an error in
Lisp (no method `hand-suit` for a list): the mirror raises the same way unless `fixed=True`, which applies the
evident intent (longest combined suit of the first and third hand).
"""
from __future__ import annotations

import numpy as np

from .cards import Hand
from .engine import Engine

SUITS = ("C", "D", "H", "S")


class LispError(RuntimeError):
    """Where the reference signals a Lisp error."""


class Outcome:
    """bridge.cl:2190-2196 (the slots a caller reads)."""

    def __init__(self, tricks, winner_nos, score):
        self.tricks, self.winner_nos, self.score = tricks, winner_nos, score

    def __repr__(self):
        return f"OUTCOME<{self.tricks} / {self.score}>"


def suitno(s):
    if s is None or isinstance(s, (int, np.integer)):
        return None if s is None else int(s)
    t = str(s).upper()
    return SUITS.index(t) if t in SUITS else None          # 'nt -> nil


def _outcome(row):
    if row["status"] == 2:
        raise LispError("play-deal: the reference signals an error on this deal")
    if row["status"] == 1:
        raise MemoryError("play-deal: the deal's route lists outgrew the per-deal arena; raise arena_kb")
    if row["status"] == 3:
        raise LispError("play-deal: internal table overflow")
    if row["status"] == 4:
        return None
    n = int(row["n_tricks"])
    tricks = [[None if c == 255 else [int(c) >> 4, int(c) & 15] for c in row["cards"][t]] for t in range(n)]
    return Outcome(tricks, [int(w) for w in row["winner"][:n]], [int(row["score"][0]), int(row["score"][1])])


def play_deal(trump, declarer: Hand, left: Hand, dummy: Hand, right: Hand, engine: Engine | None = None, arena_kb: int = 0):
    eng = engine or Engine.default()
    t = suitno(trump)
    hands = np.array([[h.mask64() for h in (declarer, left, dummy, right)]], dtype=np.uint64)
    return _outcome(eng.play_hands(hands, -1 if t is None else t, arena_kb)[0])


def best_tricks(trump, n, e, s, w, engine: Engine | None = None):
    return play_deal(trump, n, e, s, w, engine=engine).score[1]


def trump_length(trump, n, e, s, w):
    """bridge.cl:3083-3086 (takes raw suit lists in the reference; Hands accepted too)."""
    sn = suitno(trump)
    return sum(len((h.suits if isinstance(h, Hand) else h)[sn]) for h in (n, s))


def longest_suit(*hands):
    """bridge.cl:3030-3036: the suit with most cards in the first and third hand, later suit on ties."""
    if not hands or not isinstance(hands[0], Hand):
        raise LispError("no applicable method hand-suit")
    best = None
    for i, st in enumerate(SUITS):
        v = (st, len(hands[0].suits[i]) + len(hands[2].suits[i]))
        if best is None or v[1] >= best[1]:
            best = v
    return best[0]


def best_tricks_star(_, n, e, s, w, engine: Engine | None = None, fixed: bool = False):
    """bridge.cl:3038-3043."""
    trump = longest_suit(n, e, s, w) if fixed else longest_suit([n, e, s, w])
    return [trump, best_tricks(trump, n, e, s, w, engine=engine), best_tricks(None, n, e, s, w, engine=engine)]


def best_outcomes(_, n, e, s, w, engine: Engine | None = None, fixed: bool = False):
    """bridge.cl:3071-3081: for N/S and for E/W declaring, the better of the suit and the no-trump contract."""
    from .compscore import contract_score
    out = []
    hands = [n, e, s, w]
    for r in (0, 1):
        rolled = hands[-r:] + hands[:-r] if r else hands          # (roll r hands)
        suit, suit_tricks, nt_tricks = best_tricks_star(None, *rolled, engine=engine, fixed=fixed)
        if contract_score(None, nt_tricks, engine=engine) > contract_score(suit, suit_tricks, engine=engine):
            out += ["NT", nt_tricks]
        else:
            out += [suit, suit_tricks]
    return out
