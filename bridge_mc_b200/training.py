"""Host mirror of training.cl's drill set-up (training.cl:1-4, 58-79): `full-random`, `better-score` and the part of
`play-training` that is not interactive -- build an unconstrained deal, find each line's longest combined suit, have
the trick estimator evaluate the four candidate contracts (N/S in their suit, N/S in no-trump, E/W in theirs, E/W in
no-trump) and keep the best one for the drill.

The deal comes from the library (`Deal.build`, bmc_sample) and the trick estimator is the reference's `play-deal`
(bridge.cl:2961) on the GPU (`playdeal.play_deal`, bmc_play_hands).  A stand-in may be passed as
`play_deal(trump, declarer, left, dummy, right) -> outcome` (an object with `score` = (defence, declaring line), or with
a number in `tricks`).  The interactive card-by-card part (`manual-play`, reading cards from the terminal) stays where it
is."""
from __future__ import annotations

from .cards import SUIT_SYMS
from .deal import Deal, roll


def full_random(engine=None, seed=None) -> Deal:
    """training.cl:1 `(defvar full-random (make-instance 'deal))`"""
    return Deal(engine=engine, seed=seed)


def longest_suit(*hands):
    """bridge.cl:3030-3036: the suit in which the first and third hand hold most cards together; later suits win ties."""
    best = None
    for i, sym in enumerate(SUIT_SYMS):
        n = len(hands[0].suits[i]) + len(hands[2].suits[i])
        if best is None or n >= best[1]:
            best = (sym, n)
    return best[0]


def better_score(a, b) -> bool:
    """training.cl:3-4: `(> (second (score a)) (second (score b)))` -- more tricks for the declaring line"""
    def line_tricks(o):
        return o.score[1] if hasattr(o, "score") else o.tricks
    return line_tricks(a) > line_tricks(b)


def drill_candidates(deal):
    """training.cl:60-70: the four (hands, trump) candidates of a deal in the order `play-training` lists them"""
    ns = longest_suit(*deal)
    we = longest_suit(*roll(1, deal))
    return [(list(deal), ns), (list(deal), None), (roll(1, deal), we), (roll(1, deal), None)]


def play_training(play_deal=None, dealgen: Deal | None = None, engine=None):
    """training.cl:58-79 up to the point where the user starts playing: returns dict(deal, hands, trump, outcome, lead)
    for the candidate the estimator likes best (the FIRST best one: `fold` only replaces on a strictly better score).
    lead = `(first (first (tricks outcome)))`, the card the drill opens with (training.cl:75)."""
    if play_deal is None:
        from . import playdeal

        def play_deal(trump, *hands):
            return playdeal.play_deal(trump, *hands, engine=engine)
    deal = (dealgen or full_random(engine)).build()
    chosen = None
    for hands, trump in drill_candidates(deal):
        outcome = play_deal(trump, *hands)
        if chosen is None or better_score(outcome, chosen[2]):
            chosen = (hands, trump, outcome)
    hands, trump, outcome = chosen
    return {"deal": deal, "hands": hands, "dummy": hands[2], "hand": hands[0], "trump": trump, "outcome": outcome,
            "lead": getattr(outcome, "first_lead", None) if not isinstance(getattr(outcome, "tricks", None), list)
            else (outcome.tricks[0][0] if outcome.tricks else None)}
