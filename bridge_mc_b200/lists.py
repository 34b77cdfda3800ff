"""Host mirrors of the small list utilities the reference's deal path is written with
(bridge.cl:218-237, 420-441, 458-484, 548-568, 627-708, 716-791, 909-919).  Cards here are the
reference's own `(suit rank)` pairs, e.g. ["S", 12]; the library itself never sees them (it works on
bit planes), they exist so that Lisp-side callers and the reference's inline tests can be replayed.
No random numbers: `deal` / `deal-all` (bridge.cl:645-660) are what libbridgemc replaces."""
from __future__ import annotations

SUITS = ("C", "D", "H", "S")


class BadIndex(IndexError):
    """bridge.cl:621-625 condition bad-index"""

    def __init__(self, val, max_idx):
        super().__init__(f"bad index {val} / {max_idx}")
        self.val, self.max_idx = val, max_idx


def push_pos(lst, pos, element):
    """bridge.cl:218-221: cons `element` onto the list at `pos`, padding with empty lists"""
    lst = list(lst or [])
    out = []
    for i in range(max(len(lst) - 1, pos) + 1):
        cur = lst[i] if i < len(lst) else None
        out.append([element] + list(cur or []) if i == pos else cur)
    return out


def roll(amount, lst):
    """bridge.cl:420-425"""
    if not lst:
        return lst
    a = (-amount) % len(lst)
    return list(lst) if a == 0 else list(lst[a:]) + list(lst[:a])


def mod_or_push(mod, init, lst, val):
    """bridge.cl:458-463: replace the first element `mod` accepts, else append (init val)"""
    out = list(lst or [])
    for i, x in enumerate(out):
        ans = mod(x, val)
        if ans is not None and ans is not False:
            out[i] = ans
            return out
    return out + [init(val)]


def suitno(suit):
    return SUITS.index(str(suit).upper())


def suit_cmp(a, b):
    """bridge.cl:549-552 suit<>"""
    return suitno(a) - suitno(b)


def card_cmp(a, b):
    """bridge.cl:558-562 card<>"""
    ds = suit_cmp(a[0], b[0])
    return a[1] - b[1] if ds == 0 else ds


def take(idx, lst):
    """bridge.cl:627-634: (element, list without it); bad-index past the end"""
    if idx >= len(lst):
        raise BadIndex(idx, len(lst))
    return [lst[idx], list(lst[:idx]) + list(lst[idx + 1:])]


def take_match(cond, lst):
    """bridge.cl:640-642"""
    for i, x in enumerate(lst):
        if cond(x):
            return take(i, lst)
    return None


def remove_cards(cards, deck):
    """bridge.cl:662-664"""
    return [list(cards), [x for x in deck if x not in cards]]


def rank_hcp(rank):
    """bridge.cl:669-670"""
    return rank - 8 if rank > 8 else 0


def card_hcp(card):
    return rank_hcp(card[1])


def card_suitno(card):
    return suitno(card[0])


def all_cards():
    """bridge.cl:545-547: clubs first, ranks ascending"""
    return [[s, r] for s in SUITS for r in range(13)]


def cards_by(classifier, cards):
    """bridge.cl:676-681: bucket cards by a small integer class; each bucket in REVERSE order of arrival"""
    acc = []
    for c in cards:
        acc = push_pos(acc, classifier(c), c)
    return acc


def permuts(n, maximum):
    """bridge.cl:718-732: all n-tuples over 0..maximum, first position fastest, LAST generated first"""
    out = []
    cur = [0] * n
    while True:
        out.append(list(cur))
        i = 0
        while i < n and cur[i] + 1 > maximum:
            cur[i] = 0
            i += 1
        if i == n:
            break
        cur[i] += 1
    out.reverse()
    return out


def valid_hcp_sums(value, test=None, supply=None):
    """bridge.cl:734-742: (nJ nQ nK nA) count vectors whose HCP `test`s against value (default =), within supply"""
    test = test or (lambda a, b: a == b)
    out = []
    for lst in permuts(4, 4):
        hcp = sum((i + 1) * x for i, x in enumerate(lst))
        if test(hcp, value) and (supply is None or all(x <= y for x, y in zip(lst, supply))):
            out.append(lst)
    return out


def zip_deals(deals):
    """bridge.cl:786-791"""
    rest, choice = [], []
    for first, second in deals:
        rest = list(first) + rest
        choice = list(second) + choice
    return [rest, choice]


def even_sublists(size, lst):
    """bridge.cl:909-914"""
    lst = list(lst)
    return [lst[i:i + size] for i in range(0, len(lst), size)]


def divlist(sectests, lst, map=None):
    """bridge.cl:239-248: split `lst` into len(sectests) + 1 sections -- the first test an element passes, the
    last section for none; each section in REVERSE order of arrival"""
    acc = [None] * (len(sectests) + 1)
    for x in lst:
        sec = next((i for i, t in enumerate(sectests) if t(x)), len(sectests))
        acc = push_pos(acc, sec, map(x) if map else x)
    return acc


def suit_breakdown(cards):
    """bridge.cl:866-869: cards by suit, c d h and the rest (spades)"""
    return divlist([lambda c, s=s: str(c[0]).upper() == s for s in ("C", "D", "H")], cards)


def hcp_breakdown(cards):
    """bridge.cl:871-874: cards by HCP value 0 1 2 3 and the rest (aces)"""
    return divlist([lambda c, v=v: card_hcp(c) == v for v in range(4)], cards)


def supply(cards):
    """bridge.cl:876-878: per suit (n_low nJ nQ nK nA) of `cards`"""
    return [[len(x or []) for x in hcp_breakdown(suit or [])] for suit in suit_breakdown(cards)]

