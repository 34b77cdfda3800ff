"""TEST INFRASTRUCTURE -- CPU restatement of the reference's trick estimator `play-deal`
(bridge.cl:1497-2963 and its consumers 3027-3043).  Nothing under
bridge_mc_b200/ may import this file.

Pure-Python transliteration, function by function, of the reference's list code.  The
estimator is a heuristic search whose answer depends on evaluation order and on
*shared mutable state* (routes are popped in place, bridge.cl:2372; the `establish`
targets are popped in place, 2863; a play-com's guards are overwritten in place, 2614),
so the objects below are mutable Python objects used exactly where the reference uses
CLOS instances and conses, and every helper that returns a fresh list in Lisp returns
a fresh list here.

Conventions
  nil            -> None (an empty Python list also counts as nil: `L(x)`)
  suit symbols   -> 0..3 (c d h s), nil/NT -> None  (`suitno`, bridge.cl:519-522)
  card           -> [suit, rank] (rank 0..12 = 2..A), a trick's cards may hold None
  numbers        -> 0 is TRUE in Lisp: tests on ranks / indices use `is not None`

Parity status: pinned by the reference's 67 inline known-answer tests of this code
(bridge.cl:1514-2976, replayed by tests/test_playdeal_oracle.py from
tests/golden/playdeal_kats.json); the reference itself cannot be run here (no Lisp).
"""
from __future__ import annotations

import sys

sys.setrecursionlimit(10000)


class LispError(Exception):
    """Where the reference would signal an error (type error, no applicable method ...)."""


def L(x) -> bool:
    """Lisp truth: everything but nil (0 is true)."""
    return not (x is None or x is False or (isinstance(x, list) and len(x) == 0))


def nth(i, lst):
    if i is None:
        raise LispError("nth nil")
    if lst is None or i >= len(lst):
        return None
    return lst[i]


def seektree(path, tree):            # bridge.cl:210-212
    for p in path:
        if tree is not None and not isinstance(tree, list):
            raise LispError("nth on an atom")
        tree = nth(p, tree)
    return tree


def roll(amount, lst):               # bridge.cl:420-424
    if not L(lst):
        return None
    a = (-amount) % len(lst)
    if a == 0:
        return lst
    return lst[a:] + lst[:a]


def take(idx, lst):                  # bridge.cl:627-634
    if idx >= len(lst):
        raise LispError("bad-index")
    return [lst[idx], lst[:idx] + lst[idx + 1:]]


def take_match(cond, lst):           # bridge.cl:640-642
    for i, x in enumerate(lst):
        if cond(x):
            return take(i, lst)
    return None


def divlist(sectests, lst, map=lambda x: x):   # bridge.cl:239-248 (sections come out in REVERSE input order)
    acc = [[] for _ in range(len(sectests) + 1)]
    for item in lst or []:
        sec = len(sectests)
        for i, tst in enumerate(sectests):
            if L(tst(item)):
                sec = i
                break
        acc[sec].insert(0, map(item))
    return acc


def list_minus(a, b):                # bridge.cl:341-343
    return [x for x in (a or []) if x not in (b or [])]


def list_lead(n, lst):               # bridge.cl:33-35
    lst = lst or []
    return lst if len(lst) <= n else lst[:n]


def max_q(cards):                    # (apply #'max? cards)  bridge.cl:266-267
    return max(cards) if L(cards) else None


def suitno(s):                       # bridge.cl:519-522, 1528-1530
    if s is None or isinstance(s, int):
        return s
    try:
        return "CDHS".index(str(s).upper())
    except ValueError:
        return None                  # 'nt -> nil


# ---------------------------------------------------------------------------
# hand (bridge.cl:1622-1640, 1702-1724)
# ---------------------------------------------------------------------------
class Hand:
    __slots__ = ("suits", "id")

    def __init__(self, suits, id=None):
        self.suits = [sorted(s) if s else [] for s in suits]
        self.id = id if id is not None else object()     # a fresh random id string: only its identity matters (eq)

    def __repr__(self):
        return "HAND<%s>" % (self.suits,)


def mkhand(suits, id=None):          # bridge.cl:1073
    return Hand(suits, id)


def str2hand(text):                  # bridge.cl:1645-1666
    hid = None
    if ": " in text:
        hid, text = text.split(": ", 1)
    suits = [[], [], [], []]
    fields = text.split(" ")
    marks = "♣♦♥♠"
    i = 0
    while i < len(fields):
        f = fields[i]
        if f in marks and len(f) == 1:
            sn = marks.index(f)
            nxt = fields[i + 1] if i + 1 < len(fields) else ""
            if nxt in marks and len(nxt) == 1:
                suits[sn] = []
                i += 1
                continue
            ranks, j = [], 0
            while j < len(nxt):
                if nxt[j] == "1":
                    ranks.append(8)
                    j += 2
                else:
                    ranks.append("23456789_JQKA".index(nxt[j]))
                    j += 1
            suits[sn] = ranks
            i += 2
        else:
            i += 1
    return Hand(suits, hid if hid is not None else None)


def hand_suit(h, suit):              # bridge.cl:1702-1703
    if not isinstance(h, Hand):
        raise LispError("no applicable method hand-suit")
    sn = suitno(suit)
    if sn is None:
        raise LispError("nth nil")
    return h.suits[sn]


def beats(a, b, suit, trump):        # bridge.cl:1705-1714
    ac, bc = hand_suit(a, suit), hand_suit(b, suit)
    at = hand_suit(a, trump) if trump is not None else None
    bt = hand_suit(b, trump) if trump is not None else None
    if L(ac) and L(bc):
        return max(ac) > max(bc)
    if not L(ac) and not L(bc) and L(at) and L(bt):
        return max(at) > max(bt)
    if not L(ac) and L(at):
        return True
    return False


def hand_play(h, suit, chooser):     # bridge.cl:1716-1724
    if suit is None:
        raise LispError("nth nil")
    played = h.suits[suit]
    if L(played):
        ans = chooser(played)
        if L(ans):
            return [[suit, ans[0]], Hand(h.suits[:suit] + [ans[1]] + h.suits[suit + 1:], h.id)]
    return None


def play_low(suit, hand):            # bridge.cl:1500-1505 (raw suit lists)
    sn = suitno(suit)
    cards = hand[sn]
    if L(cards):
        return [cards[0], [c[1:] if i == sn else c for i, c in enumerate(hand)]]
    return None


def lowest_card(cards):              # bridge.cl:1784-1785
    return [cards[0], cards[1:]]


def closest_card(card, cards, acc=None):   # bridge.cl:1787-1790
    acc = acc or []
    if not L(cards):
        return take(len(acc) - 1, acc) if L(acc) else None
    if cards[0] >= card:
        return [cards[0], acc + cards[1:]]
    return closest_card(card, cards[1:], acc + [cards[0]])


def take_higher(cards, high, use_highest=False):   # bridge.cl:1518-1526
    if L(use_highest):
        best = [0, 0]
        for i, v in enumerate(cards):
            if v > best[1]:
                best = [i, v]
        higher = best[0] if best[1] > high else None
    else:
        higher = next((i for i, v in enumerate(cards) if high < v), None)
    return take(higher, cards) if higher is not None else None


# ---------------------------------------------------------------------------
# trick (bridge.cl:1535-1620)
# ---------------------------------------------------------------------------
class Trick:
    def __init__(self, suit=None, trump=None, high=None, winner=None, remaining=None):
        sn = suitno(suit)
        if high is None or isinstance(high, list):
            self.high = high if L(high) else [sn, -1]
        else:
            self.high = [sn, high]
        self.suit = sn
        self.trump = suitno(trump)
        self.winner = winner if winner is not None else -1
        self.cards = None
        self.remaining = remaining

    def card_higher(self, a, b):     # bridge.cl:1556-1570
        asuit, arank = nth(0, a), nth(1, a)
        bsuit, brank = nth(0, b), nth(1, b)
        atrump, as_ = asuit == self.trump, asuit == self.suit
        btrump, bs = bsuit == self.trump, bsuit == self.suit
        if atrump and not btrump:
            return True
        if btrump and not atrump:
            return False
        if as_ and not bs:
            return True
        if bs and not as_:
            return False
        if not (as_ or atrump):
            return False
        if not (bs or btrump):
            return True
        return arank > brank

    def play(self, card, rest):      # bridge.cl:1588-1599
        if L(card) and self.card_higher(card, self.high):
            self.high = card
            self.winner = len(self.remaining or [])
        if L(self.remaining):
            self.remaining.append(rest)
        else:
            self.remaining = [rest]
        if L(self.cards):
            self.cards.append(card)
        else:
            self.cards = [card]
        return self

    def winnerp(self):               # bridge.cl:1601-1605
        return self.winner >= 0 and len(self.remaining or []) % 2 == self.winner % 2

    def roll_trick(self, n):         # bridge.cl:1607-1612
        self.winner = (self.winner + n) % 4
        self.cards = roll(n, self.cards)
        self.remaining = roll(n, self.remaining)
        return self

    def score(self):                 # bridge.cl:1614-1615
        return 1 if self.winner % 2 == 0 else -1

    def as_list(self):               # bridge.cl:1726-1727
        return [self.cards[-1]] + [h.suits for h in self.remaining]

    def as_summary(self):            # bridge.cl:1729-1730
        return [self.score()] + [h.suits for h in self.remaining]


def weak_suit(hand, trick):          # bridge.cl:1732-1768
    trump, remaining = trick.trump, trick.remaining or []
    acc = None
    for suit, cards in enumerate(hand.suits):
        val = [suit, cards]
        if not L(cards):
            continue
        if acc is None:
            acc = val
            continue
        accsuit, accards = acc
        hon = any(c > 8 for c in cards)
        acchon = any(c > 8 for c in accards)
        if suit == trump:
            continue
        if accsuit == trump:
            acc = val
        elif not hon and not acchon:
            if len(cards) < len(accards):
                acc = val
        elif not hon:
            acc = val
        elif not acchon:
            pass
        else:
            rem_a = [c for h in remaining for c in hand_suit(h, suit)]
            rem_b = [c for h in remaining for c in hand_suit(h, accsuit)]
            buf_a = len(cards) - len([c for c in rem_a if max(cards) < c])
            buf_b = len(accards) - len([c for c in rem_b if max(accards) < c])
            if buf_a == buf_b:
                if cards[0] < accards[0]:
                    acc = val
            elif (buf_a > 0 and buf_b > 0) or (buf_a < 0 and buf_b < 0):
                if buf_a > buf_b:
                    acc = val
            elif buf_a > 0:
                acc = val
    return acc[0] if acc is not None else None


def beat_or_low(trick, hand, use_highest=False):   # bridge.cl:1804-1834
    suit, trump = trick.suit, trick.trump
    hsuit, hrank = nth(0, trick.high), nth(1, trick.high)

    def low_chain():
        return (hand_play(hand, weak_suit(hand, trick), lowest_card)
                or (hand_play(hand, trump, lowest_card) if trump is not None else None))

    if hsuit is not None and hsuit == trump and suit != trump:
        ans = hand_play(hand, suit, lowest_card)
        if ans is None and not trick.winnerp():
            ans = hand_play(hand, trump, lambda cards: take_match(lambda c: hrank < c, cards))
        if ans is None:
            ans = low_chain()
    elif hsuit is not None and hsuit == suit:
        ans = hand_play(hand, suit, lambda cards: ((not trick.winnerp()) and take_higher(cards, hrank, use_highest))
                        or lowest_card(cards))
        if ans is None and trump is not None:
            ans = hand_play(hand, trump, lowest_card)
        if ans is None:
            ans = low_chain()
    else:
        ans = hand_play(hand, suit, lambda cards: take_higher(cards, -1, use_highest) or lowest_card(cards))
        if ans is None and trump is not None:
            ans = hand_play(hand, trump, lowest_card)
        if ans is None:
            ans = low_chain()
    if ans is not None and L(ans[1]):
        return trick.play(ans[0], ans[1])
    return None


def finesse_if_possible(trick, h1, h2):            # bridge.cl:1905-1947
    suit, trump = trick.suit, trick.trump
    if trick.winnerp():
        hsuit = hrank = None
    else:
        hsuit, hrank = nth(0, trick.high), nth(1, trick.high)
    h2suit = nth(suit, h2.suits)

    def trump_clause():
        if trump is not None and not L(h2suit) and L(nth(trump, h2.suits)):
            h2max = max(h2.suits[trump])
            return hand_play(h1, trump, lambda cards: take_match(lambda c: h2max < c, cards))
        if trump is not None and not trick.winnerp():
            return hand_play(h1, trump, lowest_card)
        return None

    if hsuit is not None and hsuit == trump and trump != suit:
        ans = hand_play(h1, suit, lowest_card)
        if ans is None:
            if L(h2suit) and suit != trump:
                ans = hand_play(h1, trump, lambda cards: take_match(lambda c: hrank < c, cards))
            else:
                h2max = max_q(nth(trump, h2.suits))
                if h2max is None:
                    h2max = -1
                ans = hand_play(h1, trump, lambda cards: take_match(lambda c: hrank < c and h2max < c, cards))
        if ans is None:
            ans = hand_play(h1, weak_suit(h1, trick), lowest_card)
    elif hsuit is not None and hsuit == suit:
        if L(h2suit):
            h2max = max_q(h2suit)
            ans = (hand_play(h1, suit, lambda cards: take_match(lambda c: h2max < c and hrank < c, cards))
                   or hand_play(h1, suit, lambda cards: take_match(lambda c: hrank < c, cards)))
        else:
            ans = hand_play(h1, suit, lambda cards: take_match(lambda c: hrank < c, cards))
        if ans is None:
            ans = hand_play(h1, suit, lowest_card)
        if ans is None:
            ans = trump_clause()
        if ans is None:
            ans = hand_play(h1, weak_suit(h1, trick), lowest_card)
    else:
        ans = None
        if L(h2suit):
            h2max = max(h2suit)
            ans = hand_play(h1, suit, lambda cards: take_match(lambda c: h2max < c, cards))
        if ans is None:
            ans = hand_play(h1, suit, lowest_card)
        if ans is None:
            ans = trump_clause()
        if ans is None:
            ans = hand_play(h1, weak_suit(h1, trick), lowest_card)
    if ans is not None and L(ans[1]):
        return trick.play(ans[0], ans[1])
    return None


def invert_if_needed(lst):           # bridge.cl:2033-2036
    if not L(lst):
        return lst
    return [lst[0]] + (roll(2, lst[1:]) or []) if L(lst[0]) else lst


def play_round(suit, trump, dealer, left, dummy, right):   # bridge.cl:2050-2069 (raw suit lists)
    def one_side_trick(a, b, c, d):
        acc = beat_or_low(Trick(suit=suit, trump=trump), Hand(a), use_highest=True)
        for x in (b, c, d):
            if acc is None:
                raise LispError("with-slots on nil")
            acc = beat_or_low(acc, Hand(x))
        return acc

    def high_or_finesse(a, b, c, d):
        trick = Trick(suit=suit, trump=trump)
        beat_or_low(trick, Hand(a))
        beat_or_low(trick, Hand(b))
        finesse_if_possible(trick, Hand(c), Hand(d))
        beat_or_low(trick, Hand(d))
        return trick

    sn = suitno(suit)
    dummy_max, dealer_max = max_q(dummy[sn]), max_q(dealer[sn])
    if dummy_max is None and dealer_max is None:
        return None
    if dummy_max is None:
        return one_side_trick(dealer, left, dummy, right)
    if dealer_max is None:
        return one_side_trick(dummy, right, dealer, left).roll_trick(2)
    if dealer_max > dummy_max:
        return high_or_finesse(dummy, right, dealer, left).roll_trick(2)
    return high_or_finesse(dealer, left, dummy, right)


# ---------------------------------------------------------------------------
# cache, outcome (bridge.cl:2178-2267)
# ---------------------------------------------------------------------------
class Cache:                         # bridge.cl:2178-2188: `equal` on the argument list = identity of the hand objects
    def __init__(self, action):
        self.action = action
        self.hits = {}

    def run(self, trump, a, b, c, d, suits):
        key = (trump, id(a), id(b), id(c), id(d), tuple(suits or []))
        hit = self.hits.get(key)
        if hit is not None:
            return hit[0]
        ans = self.action(self, trump, a, b, c, d, suits=suits)
        self.hits[key] = (ans, a, b, c, d)      # keep the hands alive: id() must stay unique
        return ans


class Outcome:
    def __init__(self, tricks=None, winners=None, winner_nos=None, score=None, roll=0, remaining=None):
        self.tricks = tricks
        self.winners = winners
        self.winner_nos = winner_nos
        self.score = score if score is not None else [0, 0]
        self.roll = roll
        self.remaining = remaining

    def __repr__(self):
        return "OUTCOME<%s / %s>" % (self.tricks, self.score)


def new_outcome(x):                  # bridge.cl:2204-2210, 2237-2240
    if isinstance(x, list):
        acc = new_outcome(x[0])
        for t in x[1:]:
            acc = add_outcome_star(acc, new_outcome(t))
        return acc
    trick = x
    if trick.winner < 0:
        raise LispError("nth -1")
    return Outcome(tricks=[trick.cards], winners=[nth(trick.winner, trick.cards)], winner_nos=[trick.winner],
                   score=[1, 0] if trick.winner % 2 == 0 else [0, 1], roll=trick.winner,
                   remaining=roll(-trick.winner, trick.remaining))


def _plus_score(self, other):
    b = list(reversed(other.score)) if self.roll % 2 == 1 else other.score
    return [x + y for x, y in zip(self.score, b)]


def add_outcome(self, other):        # bridge.cl:2212-2223 (mutates self)
    self.tricks = (self.tricks or []) + (other.tricks or [])
    self.winners = (self.winners or []) + (other.winners or [])
    self.winner_nos = (self.winner_nos or []) + (other.winner_nos or [])
    self.score = _plus_score(self, other)
    self.roll = other.roll
    self.remaining = other.remaining
    return self


def add_outcome_star(self, other):   # bridge.cl:2225-2235 (fresh instance)
    return Outcome(tricks=(self.tricks or []) + (other.tricks or []), winners=(self.winners or []) + (other.winners or []),
                   winner_nos=(self.winner_nos or []) + (other.winner_nos or []), score=_plus_score(self, other),
                   roll=(self.roll + other.roll) % 4, remaining=other.remaining)


def won(self):                       # bridge.cl:2242-2244
    return L(self.winner_nos) and self.winner_nos[0] % 2 == 0


def do_next(self, if_won, if_lost=None):   # bridge.cl:2246-2251
    if L(self.winner_nos):
        if won(self):
            return add_outcome(self, if_won(*self.remaining))
        if if_lost is not None:
            return add_outcome(self, if_lost(*self.remaining))
    return None


def best_outcome(*outcomes):         # bridge.cl:2253-2258
    acc = outcomes[0] if outcomes else None
    for val in outcomes[1:]:
        if acc is None or (val is not None and acc.score[0] < val.score[0]):
            acc = val
    return acc


def best_outcome_star(rollv, *outcomes):   # bridge.cl:2260-2267
    cur = rollv % 2
    acc = outcomes[0] if outcomes else None
    for val in outcomes[1:]:
        if acc is None or (val is not None and acc.score[cur] < val.score[cur]):
            acc = val
    return acc


def simple_trick(trump, suit, a, b, c, d, use_highest=False):   # bridge.cl:2269-2277
    if len(hand_suit(a, suit)) > 0:
        trick = Trick(suit=suit, trump=trump)
        beat_or_low(trick, a, use_highest=use_highest)
        if beats(d, c, suit, trump):
            beat_or_low(trick, b)
        else:
            finesse_if_possible(trick, b, c)
        finesse_if_possible(trick, c, d)
        beat_or_low(trick, d)
        return new_outcome(trick)
    return Outcome(remaining=[a, b, c, d])


def _four_ways(hands):               # (list-prod '(T NIL) (list hands (roll 2 hands)))  bridge.cl:2288, 2482
    return [(True, hands), (True, roll(2, hands)), (False, hands), (False, roll(2, hands))]


def suit_tricks(trump, suit, a, b, c, d):      # bridge.cl:2279-2296
    def play_through(*hands):
        hands = list(hands)
        outs = []
        for top, hs in _four_ways(hands):
            o = do_next(simple_trick(trump, suit, *hs, use_highest=top), play_through, play_through)
            outs.append(o if o is not None else Outcome(remaining=hs))
        return best_outcome(*outs)

    out = play_through(a, b, c, d)
    seq = [[player % 2 == 0, card] for player, card in zip(out.winner_nos or [], out.winners or [])]
    acc = None
    for first, card in reversed(seq):
        car = acc[0] if L(acc) else None
        cdr = acc[1:] if L(acc) else []
        if first:
            acc = [[card] + (car or [])] + cdr
        else:
            acc = [None, [card] + (car or [])] + cdr
    return acc


class SuitValue:                     # bridge.cl:2298-2327
    def __init__(self, suit, deal, trump=None):
        self.suit = suit
        self.trump = trump
        self.tricks = suit_tricks(trump, suit, *deal)
        self.summary = [len(t or []) for t in (self.tricks or [])]

    def ours(self):
        return sum(self.summary[0::2])

    def theirs(self):
        return sum(self.summary[1::2])

    def better(self, other):
        return other if (other is not None and other.ours() > self.ours()) else self


def deal_plan_suits(deal, trump):    # bridge.cl:2336-2356 (`suits` slot; iwins / ilosers are not read by the estimator)
    trump = suitno(trump)
    vals = []
    for suit in range(4):
        nt = SuitValue(suit, deal)
        vals.append(nt.better(SuitValue(suit, deal, trump) if trump is not None else None))
    vals = [v for v in vals if 0 < v.ours()]
    vals.sort(key=lambda v: -v.ours())          # stable, like SBCL's merge sort of lists
    if trump is not None:
        return ([v for v in vals if v.trump is not None] + [v for v in vals if v.suit == trump]
                + [v for v in vals if not (trump == v.suit or v.trump is not None)])
    return vals


def find_opp_weakness(trump, a, b, c, d):      # bridge.cl:2745-2759
    out = []
    for sv in deal_plan_suits([a, b, c, d], trump):
        tr = sv.tricks or []
        ours = [cd[1] for i in range(0, len(tr), 2) for cd in (tr[i] or []) if cd[0] == sv.suit]
        theirs = [cd[1] for i in range(1, len(tr), 2) for cd in (tr[i] or []) if cd[0] == sv.suit]
        out.append([sv.suit, sv.ours() - sv.theirs(), ours, theirs])
    return out


# ---------------------------------------------------------------------------
# routes (bridge.cl:2358-2660)
# ---------------------------------------------------------------------------
class PlayRoute:
    def __init__(self, to, tricks, frm=None, to_suit_void=None):
        self.frm = frm
        self.to = to
        self.to_suit_void = to_suit_void
        self.tricks = tricks

    def __repr__(self):
        return "PLAY-ROUTE<%s->%s @ %s: %s>" % (self.frm, self.to, self.to_suit_void, self.tricks)


def reproduce(route, hands, trump):            # bridge.cl:2369-2395
    while True:
        if not L(route.tricks):
            return None
        base = route.tricks[0]
        route.tricks = route.tricks[1:]         # (pop tricks): the route object itself forgets the trick
        c0 = nth(0, base)
        if c0 is None:
            raise LispError("nth nil")
        ref = base if nth(1, c0) in hand_suit(hands[0], nth(0, c0)) else roll(2, base)
        suit = seektree((0, 0), ref)
        rank0 = seektree((0, 1), ref)
        first = hand_play(hands[0], suit, lambda cards: closest_card(rank0, cards))
        ans = None
        if first is not None:
            trick = Trick(suit=suit, trump=trump)
            trick.play(first[0], first[1])
            if beats(hands[3], hands[2], suit, trump):
                beat_or_low(trick, hands[1])
            else:
                finesse_if_possible(trick, hands[1], hands[2])
            s2, r2 = seektree((2, 0), ref), seektree((2, 1), ref)
            third = hand_play(hands[2], s2, lambda cards: closest_card(r2, cards))
            if third is None:
                third = hand_play(hands[2], weak_suit(hands[2], trick), lowest_card)
            if third is None:
                raise LispError("play: missing arguments")
            trick.play(third[0], third[1])
            beat_or_low(trick, hands[3])
            ans = new_outcome(trick)
        if ans is not None and won(ans):
            return ans


def rem_break_div(route, trump, cards):        # bridge.cl:2397-2411
    tricks = route.tricks or []
    break_cards = [nth(2, t) for t in tricks] if route.frm is None else None
    if route.frm is None or route.frm == route.to:
        remove = [[t[0]] if L(nth(0, t)) else None for t in tricks]
    else:
        remove = [[nth(0, t), nth(2, t)] for t in tricks]
    items = []
    for card in cards:
        bid = break_cards.index(card) if (break_cards and card in break_cards) else None
        rid = next((i for i, rc in enumerate(remove) if rc and card in rc), None)
        items.append([card, bid, rid])
    return divlist([lambda it: it[2] is not None, lambda it: it[1] is not None], items,
                   map=lambda it: [it[0], it[1] if it[1] is not None else it[2]])


def del_at(route, ids):              # bridge.cl:2413-2419
    return PlayRoute(route.to, [t for i, t in enumerate(route.tricks or []) if i not in ids], route.frm, route.to_suit_void)


def add_at(route, new_tricks):       # bridge.cl:2421-2425
    return PlayRoute(route.to, (route.tricks or []) + (new_tricks or []), route.frm, route.to_suit_void)


class PlayCom:                       # bridge.cl:2427-2451
    def __init__(self, zero_id, trump, active, pending, guards=None):
        self.zero_id = zero_id
        self.trump = trump
        self.active = active
        self.pending = pending
        self.guards = guards


def mk_play_com(trump, routes, guards, a, c):
    def is_active(route):
        v = route.to_suit_void
        return v is None or not (L(hand_suit(a, v)) and L(hand_suit(c, v)))

    _empty, active, pending = divlist([lambda r: not L(r.tricks), is_active], routes)
    return PlayCom(a.id, trump, active, pending, guards)


def _route_sort(routes):
    """(sort routes pred) with the reference's predicate (bridge.cl:2458-2462), which is not a strict weak order:
    routes with from = to precede everything (also one another), routes with a `from` precede those without (also one
    another).  A merge sort that moves the later element forward iff (pred later earlier) -- SBCL's list sort --
    yields: from = to routes in reverse input order, then other `from` routes in reverse input order, then the rest
    in input order."""
    x = [r for r in routes if r.frm is not None and r.frm == r.to]
    y = [r for r in routes if r.frm is not None and r.frm != r.to]
    z = [r for r in routes if r.frm is None]
    return x[::-1] + y[::-1] + z


def mk_route(com, a, b, c, d):       # bridge.cl:2453-2474
    def player_routes(player):
        return _route_sort([r for r in (com.active or []) if r.frm is None or r.frm == player])

    acc = Outcome(remaining=[a, b, c, d])
    pr = [player_routes(0), player_routes(2)]
    while L(pr[0]):
        ans = reproduce(pr[0][0], acc.remaining, com.trump)
        if ans is None:
            pr = [pr[0][1:], pr[1]]
        elif ans.winner_nos[0] == 0:
            acc = add_outcome(acc, ans)
        else:
            acc = add_outcome(acc, ans)
            pr = [pr[1], pr[0]]
    return acc


def play_routes(trump, a, b, c, d, suits=(0, 1, 2, 3)):   # bridge.cl:2476-2551
    suits = list(suits)

    def suit_tops(suit, *hands):
        hands = list(hands)
        outs = [do_next(simple_trick(trump, suit, *hs, use_highest=top), lambda *r: suit_tops(suit, *r))
                for top, hs in _four_ways(hands)]
        return best_outcome(*outs) or Outcome(remaining=hands)

    def guards_of(suit, *hands):
        top = sorted(([hid, r] for hid, h in enumerate(hands) for r in hand_suit(h, suit)), key=lambda e: -e[1])
        last = next((i for i, e in enumerate(top) if e[0] % 2 == 1), None)
        if last is None:
            return [e[1] for e in top]
        if last == 0:
            return None
        return [e[1] for e in top[:last]]

    def trick_type(left=False, winnerno=None, my=False, partners=False):
        def pred(item):
            cards, winner, mine = nth(0, item), nth(1, item), nth(2, item)
            if left and seektree((0, 0), cards) == seektree((2, 0), cards):
                return False
            if winnerno is not None and winner != winnerno:
                return False
            if my:
                return mine is not None
            if partners:
                return mine is None
            return True
        return pred

    def void_vertices(frm, to, tricks):
        secs = divlist([(lambda t, s=s: (nth(0, nth(0, t)) == s) if nth(0, t) is not None else None) for s in (0, 1, 2)], tricks)
        return [PlayRoute(to, st, frm, suit) for suit, st in zip(suits, secs) if L(st)]

    items = []
    for suit in suits:
        tops = suit_tops(suit, a, b, c, d)
        g = guards_of(suit, *roll(1, tops.remaining))
        for cards, wno, winner in zip(tops.tricks or [], tops.winner_nos or [], tops.winners or []):
            mine = winner[1] if winner[1] in hand_suit(a, winner[0]) else None
            items.append([cards, wno, mine])
        if L(g):
            items.append([[suitno(suit), g], 1])
    (guards, my_left, partner_left, my_trumped, partner_trumped, my_lead, partner_lead, my, partner) = divlist(
        [trick_type(winnerno=1), trick_type(left=True, winnerno=0, my=True), trick_type(left=True, winnerno=0, partners=True),
         trick_type(left=True, winnerno=2, partners=True), trick_type(left=True, winnerno=2, my=True),
         trick_type(winnerno=2, partners=True), trick_type(winnerno=2, my=True), trick_type(my=True)],
        items, map=lambda it: it[0])
    routes = ([PlayRoute(0, my), PlayRoute(2, partner), PlayRoute(2, my_lead, frm=0), PlayRoute(0, partner_lead, frm=2)]
              + void_vertices(0, 0, my_left) + void_vertices(0, 2, my_trumped)
              + void_vertices(2, 2, partner_left) + void_vertices(2, 0, partner_trumped))
    return mk_play_com(trump, routes, guards, a, c)


def insert_route(route, routes):     # bridge.cl:2553-2561 (the void suit is compared with itself: always equal)
    routes = routes or []
    for p, r in enumerate(routes):
        if r.frm == route.frm and r.to == route.to:
            return routes[:p] + [add_at(r, route.tricks)] + routes[p + 1:]
    return [route] + routes


def append_routes(a, *other):        # bridge.cl:2563-2566
    acc = a
    for lst in other:
        for val in lst or []:
            acc = insert_route(val, acc)
    return acc


def first_pass(com, cards):          # bridge.cl:2568-2590
    routes, remaining = None, cards
    for route in com.active or []:
        if L(remaining):
            remove, brk, absent = rem_break_div(route, com.trump, remaining)
            ids = [x[1] for x in remove + brk]
            updated = del_at(route, ids) if L(ids) else route
            break_ids = list_minus([x[1] for x in brk], [x[1] for x in remove])
            if L(break_ids):
                other = insert_route(PlayRoute(route.to, [nth(i, route.tricks) for i in break_ids], frm=route.to), routes)
            else:
                other = routes
            routes = ([updated] + (other or [])) if L(updated.tricks) else other
            remaining = [x[0] for x in absent]
        else:
            routes = [route] + (routes or [])
            remaining = None
    return routes, remaining


def second_pass(com, hands):         # bridge.cl:2592-2596
    return divlist([lambda r: not L(nth(suitno(r.to_suit_void), hands[r.to].suits))], com.pending)


def broken_guards(com, cards):       # bridge.cl:2598-2615 (overwrites com.guards)
    remaining, broken = [], []
    for suit, ranks in com.guards or []:
        played = [cd[1] for cd in (cards or []) if cd is not None and cd[0] == suit]
        new_ranks = list_minus(ranks, played)
        if L(new_ranks):
            remaining = [[suit, new_ranks]] + remaining
        else:
            broken = [suit] + broken
    com.guards = remaining
    return broken


def normalize_hands(com, outcome):   # bridge.cl:2622-2638
    offset = next((i for i, h in enumerate(outcome.remaining) if h.id is com.zero_id), None)
    if offset is None:
        raise LispError("(> nil 0)")

    def norm(routes):
        return [PlayRoute((r.to + offset) % 4, r.tricks, (r.frm + offset) % 4 if r.frm is not None else None, r.to_suit_void)
                for r in routes or []]

    if offset > 0:
        return PlayCom(outcome.remaining[0].id, com.trump, norm(com.active), norm(com.pending))
    return com


def com_after_tricks(com, outcome):  # bridge.cl:2643-2660
    new_self = normalize_hands(com, outcome)
    raw = [cd for t in (outcome.tricks or []) for cd in t]
    base_routes, skipped = first_pass(new_self, raw)
    new_active, new_pending = second_pass(new_self, outcome.remaining)
    broken = broken_guards(new_self, skipped)
    refreshed = play_routes(new_self.trump, *outcome.remaining, suits=broken) if L(broken) else None
    return PlayCom(new_self.zero_id, com.trump,
                   append_routes(list(reversed(base_routes or [])), new_active, refreshed.active if refreshed else None),
                   append_routes(new_pending, new_self.pending, refreshed.pending if refreshed else None),
                   (new_self.guards or []) + ((refreshed.guards or []) if refreshed else []))


# ---------------------------------------------------------------------------
# strategies and the search (bridge.cl:2705-2963)
# ---------------------------------------------------------------------------
def immediate_tricks(cache, trump, a, b, c, d, suits=(0, 1, 2, 3)):   # bridge.cl:2705-2722
    suits = list(suits or [])
    tn = suitno(trump)
    cands = []
    for suit in suits:
        cards = hand_suit(a, suit)
        if suitno(suit) == tn:
            opts = None
            if L(cards) and (L(hand_suit(b, trump)) or L(hand_suit(d, trump))):
                opts = [None, True] if len(cards) > 1 else [None]
        elif L(cards):
            opts = [None] if len(cards) < 2 else [None, True]
        else:
            opts = None
        cands += [[suit, o] for o in (opts or [])]
    acc = Outcome(remaining=[a, b, c, d])
    for suit, top in cands:
        acc = best_outcome(acc, do_next(simple_trick(trump, suit, a, b, c, d, use_highest=top),
                                        lambda *r: cache.run(trump, *r, suits=suits)))
    return acc


def hunt_card(trump, card, a, b, c, d):        # bridge.cl:2788-2812
    suit = card[0]
    trick = Trick(suit=suit, trump=trump)
    if card[1] in hand_suit(b, suit):
        tops = (False, True, False, False)
    elif card[1] in hand_suit(d, suit):
        tops = (True, False, False, False) if beats(b, c, suit, trump) else (False, False, True, False)
    else:
        tops = (False, False, False, False)
    for h, top in zip((a, b, c, d), tops):
        beat_or_low(trick, h, use_highest=top)
    return new_outcome(trick)


class PlayStrategy:                  # bridge.cl:2821-2845
    def __init__(self, trump=None, hands=None, soonest=None, establish=None, routes=None):
        self.trump = trump
        if hands is None:
            self.soonest, self.establish, self.routes = soonest, establish, routes
        else:
            targets = find_opp_weakness(trump, *hands)
            self.soonest = [t[0] for t in targets if t[1] == 0 and not L(t[2])]
            self.establish = [t for t in targets if 0 < t[1]]
            self.routes = play_routes(trump, *hands)

    def after_tricks(self, outcome):
        return PlayStrategy(trump=self.trump, soonest=self.soonest, establish=self.establish,
                            routes=com_after_tricks(self.routes, outcome))


def play_strategy(strategy, cache, a, b, c, d):           # bridge.cl:2847-2866
    trump = strategy.trump
    tn = suitno(trump)
    suits = [s for s in (strategy.soonest or []) if L(hand_suit(a, s))]
    out = cache.run(trump, a, b, c, d, suits=list_lead(2, suits))
    if L(out.tricks):
        return out

    def ok(suit):
        return (L(hand_suit(a, suit)) and (suit != tn or L(hand_suit(b, suit)) or L(hand_suit(d, suit)))
                and (trump is None or tn == suit or (L(hand_suit(b, suit)) and L(hand_suit(d, suit)))))

    target = next((t for t in (strategy.establish or []) if t[0] is not None and ok(t[0])), None)
    suit = target[0] if target is not None else None
    if suit is not None:
        out = immediate_tricks(cache, trump, a, b, c, d, suits=[suit] + list(strategy.soonest or []))
        if L(out.tricks):
            return out
        rank = None
        if L(target[3]):
            rank = target[3][0]
            target[3] = target[3][1:]           # (pop (fourth target)): shared by every strategy derived from this one
        if rank is not None:
            return hunt_card(trump, [suit, rank], a, b, c, d)
        return simple_trick(trump, suit, a, b, c, d)
    return Outcome(remaining=[a, b, c, d])


class DealPlay:                      # bridge.cl:2912-2959
    def __init__(self, hands, trump, cache=None, active=None, passive=None, outcome=None):
        self.hands = hands if active is not None else roll(-1, hands)
        self.trump = trump
        self.outcome = outcome
        self.cache = cache or Cache(immediate_tricks)
        self.active = active or PlayStrategy(trump=trump, hands=roll(-1, hands))
        self.passive = passive or PlayStrategy(trump=trump, hands=hands)

    def after_action(self, action):
        if callable(action):
            fn = action
        elif isinstance(action, list) and action and action[0] == "HIGH":
            fn = lambda n, e, s, w: simple_trick(self.trump, action[1], n, e, s, w, use_highest=True)
        elif action == "PLAY-STRATEGY":
            fn = lambda *h: play_strategy(self.active, self.cache, *h)
        elif action == "MK-ROUTE":
            fn = lambda *h: mk_route(self.active.routes, *h)
        else:
            fn = lambda *h: simple_trick(self.trump, action, *h)
        ans = fn(*self.hands)
        if L(ans.tricks):
            new_a = self.active.after_tricks(ans)
            new_p = self.passive.after_tricks(ans)
            w = won(ans)
            return DealPlay(ans.remaining, self.trump, outcome=add_outcome_star(self.outcome, ans) if self.outcome else ans,
                            active=new_a if w else new_p, passive=new_p if w else new_a)
        return None

    def play_outcome(self, stats=None):
        any_suit = next((i for i, s in enumerate(self.hands[0].suits) if L(s)), None)
        rollv = self.outcome.roll if self.outcome else 0
        if stats is not None:
            stats["nodes"] = stats.get("nodes", 0) + 1
        if any_suit is None:
            return self.outcome
        options = []
        for action in ("PLAY-STRATEGY", "MK-ROUTE", any_suit):
            ans = self.after_action(action)
            options.append(ans.play_outcome(stats) if ans is not None else None)
        return best_outcome_star(rollv, *options)


def play_deal(trump, declarer, left, dummy, right, stats=None):   # bridge.cl:2961-2963
    t = None if (trump is None or str(trump).upper() == "NT") else trump
    return DealPlay([declarer, left, dummy, right], t).play_outcome(stats)


def best_tricks(trump, n, e, s, w):  # bridge.cl:3027-3028
    return play_deal(trump, n, e, s, w).score[1]


def trump_length(trump, n, e, s, w):  # bridge.cl:3083-3086 (raw suit lists)
    return len(n[suitno(trump)]) + len(s[suitno(trump)])
