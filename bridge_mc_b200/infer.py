"""Host mirror of the reference's constraint narrowing (bridge.cl:1085-1281):
infer-hcp-range, infer-suit-ranges, infer-distgen-params and the `bad-range`
condition.  Host-side logic, as in the reference (it runs once per `deal`
object here, not once per deal).

A def is a dict in plist order with lower-case keys 'hcp' 'c' 'd' 'h' 's';
a value is a number, [lo, hi] or -- for suits -- [lo, hi, suit-hcp-range];
None = nil.  `supply` is the reference's per-suit (n_low nJ nQ nK nA) list.
"""
from __future__ import annotations

from fractions import Fraction

from ._lib import BadRange

SUIT_KEYS = ("c", "d", "h", "s")


class BadRangeError(BadRange):
    """bridge.cl:1109-1123 `bad-range` raised by the host-side narrowing."""

    def __init__(self, msg, down=None, up=None, val=None):
        RuntimeError.__init__(self, f"{msg}: " + (f"val '{val}' in <{down}-{up}>" if val is not None else f"<{down}-{up}>"))
        self.code = -3
        self.down, self.up, self.val = down, up, val


def plist_to_def(plist):
    """[:HCP (11 15) :C (6 nil)] (as read by sexp) -> {'hcp': [11, 15], 'c': [6, None]}"""
    if plist is None:
        return {}
    if isinstance(plist, dict):
        return plist
    out = {}
    for i in range(0, len(plist) - 1, 2):
        out[str(plist[i]).lstrip(":").lower()] = plist[i + 1]
    return out


def supply(cards):
    """bridge.cl:876-878: cards = iterable of card ids (13*suit+rank) -> 4 x [n_low, nJ, nQ, nK, nA]"""
    out = [[0, 0, 0, 0, 0] for _ in range(4)]
    for c in cards:
        s, r = divmod(int(c), 13)
        out[s][0 if r <= 8 else r - 8] += 1
    return out


FULL_SUPPLY = [[9, 1, 1, 1, 1] for _ in range(4)]


def _minus_q(a, b):
    """bridge.cl:1101-1104 -?"""
    if a is None:
        return b
    if b is None:
        return a
    return a - b


def _listp(x):
    return x is None or isinstance(x, (list, tuple))


def _hands_in_supply(sup):
    return Fraction(sum(sum(s) for s in sup), 13)


def _total_hcp(sup):
    return sum(h * n for s in sup for h, n in zip((1, 2, 3, 4), s[1:]))


def infer_hcp_range(sup, base, *others):
    """bridge.cl:1126-1153 -> {} or {'hcp': ...}"""
    base = plist_to_def(base)
    others = [plist_to_def(o) for o in others]
    total = _total_hcp(sup)
    if _hands_in_supply(sup) == len(others) + 1:
        acc = total
        for d in others:
            hcp = d.get("hcp")
            if hcp is None:
                acc = 0                                   # :default 0
            elif acc != 0:
                acc = max(_minus_q(acc, (hcp[1] if len(hcp) > 1 else None) if _listp(hcp) else hcp), 0)
        min_hcp = acc
    else:
        min_hcp = 0
    max_hcp = total
    for d in others:
        hcp = d.get("hcp")
        if hcp is not None:
            max_hcp = _minus_q(max_hcp, hcp[0] if _listp(hcp) else hcp)
    hcp = base.get("hcp")
    up_limit = min(max_hcp, total)
    if hcp is None:
        if up_limit < total or min_hcp > 0:
            return {"hcp": [min_hcp, up_limit]}
        return {}
    if _listp(hcp):
        down = max(hcp[0] if hcp[0] is not None else 0, min_hcp)
        up = min(hcp[1], up_limit) if (len(hcp) > 1 and hcp[1] is not None) else up_limit
        if down > up:
            raise BadRangeError("Can't infer correct limit!", down, up)
        return {"hcp": [down, up]}
    if hcp > up_limit or hcp < min_hcp:
        raise BadRangeError("Can't infer correct limit!", min_hcp, max_hcp, hcp)
    return {"hcp": hcp}


def infer_suit_ranges(sup, base, *others):
    """bridge.cl:1192-1234 -> dict of narrowed suit specs in c d h s order"""
    base = plist_to_def(base)
    others = [plist_to_def(o) for o in others]
    suitlen = [sum(s) for s in sup]
    while len(suitlen) < 4:
        suitlen.append(0)
    if _hands_in_supply(sup) == len(others) + 1:
        acc = list(suitlen)
        for d in others:
            nxt = []
            for key, a in zip(SUIT_KEYS, acc):
                val = d.get(key)
                if val is None:
                    nxt.append(0)                         # :default 0
                elif a == 0:
                    nxt.append(a)
                else:
                    if _listp(val):
                        sub = val[1] if (len(val) > 1 and val[1] is not None) else val[0]
                    else:
                        sub = val
                    nxt.append(_minus_q(a, sub))
            acc = nxt
        min_len = acc
    else:
        min_len = [0] * len(suitlen)
    max_len = list(suitlen)
    for d in others:
        nxt = []
        for key, a in zip(SUIT_KEYS, max_len):
            val = d.get(key)
            nxt.append(a if val is None else _minus_q(a, val[0] if _listp(val) else val))
        max_len = nxt
    out = {}
    for i, key in enumerate(SUIT_KEYS):
        if i >= len(sup):
            break
        slen, maxlen, minlen = suitlen[i], max_len[i], min_len[i]
        # (:hcp third-element) of every other def, nil where the chain breaks
        suit_hcp = []
        for d in others:
            v = d.get(key)
            third = None
            if v is not None:
                lst = v if _listp(v) else [v]
                third = lst[2] if len(lst) > 2 else None
            suit_hcp.append({"hcp": third} if third is not None else {})
        spec = base.get(key)
        up_limit = min(maxlen, slen)
        if spec is None:
            if up_limit < slen:
                out[key] = [minlen, up_limit]
        elif _listp(spec):
            down = max(spec[0], minlen) if spec[0] is not None else minlen
            up = min(spec[1], up_limit) if (len(spec) > 1 and spec[1] is not None) else up_limit
            hcp = spec[2] if len(spec) > 2 else None
            if down > up_limit:
                raise BadRangeError("Can't infer correct limit!", down, up)
            val = [down, up]
            if suit_hcp:                                   # non-empty list, even of nils
                r = infer_hcp_range([sup[i]], {"hcp": hcp} if hcp is not None else {}, *suit_hcp)
                if r:
                    val.append(r["hcp"])
            out[key] = val
        else:
            if spec > up_limit or spec < minlen:
                raise BadRangeError("Can't infer correct limit!", minlen, up_limit, spec)
            out[key] = spec
    return out


def infer_distgen_params(sup, base, others):
    """bridge.cl:1267-1269"""
    out = dict(infer_hcp_range(sup, base, *others))
    out.update(infer_suit_ranges(sup, base, *others))
    return out
