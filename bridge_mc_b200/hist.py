"""Host mirror of the reference's tally pipeline (bridge.cl:1325-1436):
hist-base -> hist-breakdown -> hist-percent (= `histogram`) and print-hist.

Counting (hist-base) is what the GPU replaces: the device tallies of
`bmc_sample` are handed to `hist_base` as its pre-counted `base` argument
(bridge.cl:1325-1330, KAT 1340-1342) through `tallies_as_base`; the prefix tree
and the exact ratios stay on the host, as in the reference.

Values: a Lisp list is a Python list, nil is None (an empty list is also nil),
ratios are fractions.Fraction.
"""
from __future__ import annotations

from fractions import Fraction


def _nil(x):
    return x is None or (isinstance(x, list) and len(x) == 0)


def _listp(x):
    return x is None or isinstance(x, list)


def _ensure_list(x):
    """bridge.cl:37-38"""
    if x is None:
        return []
    return x if isinstance(x, list) else [x]


def _car_star(x):
    """bridge.cl:27-28 car*: the element of a one-element list, else x (nil stays nil)"""
    if isinstance(x, list):
        if len(x) == 1:
            return x[0]
        if len(x) == 0:
            return None
    return x


def _nth(i, lst):
    lst = _ensure_list(lst) if _listp(lst) else lst
    return lst[i] if i < len(lst) else None


def _eq(a, b):
    """`equal` with nil == ()"""
    if _nil(a) and _nil(b):
        return True
    return a == b


def mod_or_push(mod, init, lst, val):
    """bridge.cl:460-465: replace the first element for which (mod elem val) is
    non-nil by that result, else append (init val)."""
    lst = list(lst or [])
    for i, ref in enumerate(lst):
        ans = mod(ref, val)
        if not _nil(ans):
            return lst[:i] + [ans] + lst[i + 1:]
    return lst + [init(val)]


def getfork(a, b):
    """bridge.cl:492-496: (common-prefix rest-a rest-b)"""
    a, b = list(a or []), list(b or [])
    i = 0
    while i < len(a) and i < len(b) and _eq(a[i], b[i]):
        i += 1
    return a[:i], a[i:], b[i:]


def hist_base(vals, base=None):
    """bridge.cl:1325-1330: count equal values in first-seen order; `base` holds
    pre-counted ((value count) ...) rows that are extended."""
    def histmod(ref, val):
        if _eq(ref[0], val):
            return [val, ref[1] + 1]
        return None

    acc = [list(r) for r in (base or [])]
    for v in vals:
        acc = mod_or_push(histmod, lambda x: [x, 1], acc, v)
    return acc


def hist_breakdown(histogram):
    """bridge.cl:1344-1373: merge rows that share a leading value into a prefix tree."""
    def rebase(newbase, basefreq, vals):
        if not _nil(newbase) and not _nil(vals):
            return [[_ensure_list(newbase) + _ensure_list(x[0]), x[1]] for x in vals]
        if not _nil(newbase):
            return [[_car_star(newbase), basefreq]]
        return vals

    def brkmod(a, b):
        if _nil(a):
            return None
        aval, afreq, achildren = _nth(0, a), _nth(1, a), _nth(2, a)
        bval, bfreq = _nth(0, b), _nth(1, b)
        common, arest, brest = getfork(_ensure_list(aval), _ensure_list(bval))
        if not common:
            return None
        if not (arest or brest):
            return [aval, afreq + bfreq]
        if len(arest) == 0 and _nil(achildren):
            return [_car_star(common), afreq + bfreq, [[None, afreq], [_car_star(brest), bfreq]]]
        return [_car_star(common), afreq + bfreq,
                mod_or_push(brkmod, lambda x: x, rebase(arest, afreq, achildren), [_car_star(brest), bfreq])]

    acc = []
    for row in histogram:
        acc = mod_or_push(brkmod, lambda x: x, acc, row)
    return acc


def hist_percent(histogram):
    """bridge.cl:1390-1398: exact ratios, children appended after (val ratio), sorted by ratio descending."""
    total = sum(_nth(1, e) for e in histogram)
    out = []
    for entry in histogram:
        val, freq, children = _nth(0, entry), _nth(1, entry), _nth(2, entry)
        head = [val, Fraction(freq, total)]
        out.append(head + hist_percent(children) if not _nil(children) else head)
    out.sort(key=lambda e: -e[1])          # list sort in SBCL is a stable merge sort
    return out


def histogram(data):
    """bridge.cl:1411-1412"""
    return hist_percent(hist_breakdown(hist_base(data)))


def _show(v):
    from . import sexp
    try:
        return sexp.dumps(v)
    except TypeError:
        return str(v)


def print_hist(hist, indent=0, out=None):
    """bridge.cl:1429-1436: one line per entry, ratios below 1/1000 as 1/n, else ~,2f%."""
    lines = [] if out is None else out
    for entry in hist:
        val, freq = entry[0], entry[1]
        if freq < Fraction(1, 1000):
            shown = Fraction(1, freq.denominator // freq.numerator)
            text = f"{shown.numerator}/{shown.denominator}"
        else:
            text = f"{float(freq * 100):.2f}%"
        lines.append("    " * indent + f"{_show(val)}: {text}")
        if len(entry) > 2:
            print_hist(entry[2:], indent + 1, lines)
    if out is None:
        print("\n".join(lines))
    return lines


def tallies_as_base(counts, label=None):
    """Device tally vector -> hist-base `base` rows ((value count) ...), zero bins
    dropped.  `label(i)` maps a bin index to the tallied value (default: i)."""
    rows = []
    for i, c in enumerate(counts):
        c = int(c)
        if c:
            rows.append([label(i) if label else i, c])
    return rows
