"""Helpers to replay the reference's inline `(test FORM EXPECTED PRED)` known-answer
tests from tests/golden/kats.json (made by tests/golden/make_golden.py)."""
import json
import os

from bridge_mc_b200 import sexp
from bridge_mc_b200.cards import str2hand

HERE = os.path.dirname(os.path.abspath(__file__))


def load_kats(head=None, file=None):
    kats = json.load(open(os.path.join(HERE, "golden", "kats.json"), encoding="utf-8"))
    out = []
    for k in kats:
        form = sexp.read(k["form"])
        if head and form[0] != head:
            continue
        if file and k["file"] != file:
            continue
        out.append((f"{k['file']}:{k['line']}", form, unquote(sexp.read(k["expected"]))))
    return out


def unquote(x):
    """'(a b) -> (a b)"""
    if isinstance(x, list) and len(x) == 2 and x[0] == "QUOTE":
        return x[1]
    return x


def hand_of(form):
    """(STR2HAND "...") -> Hand"""
    assert form[0] == "STR2HAND", form
    return str2hand(str(form[1]))


def bid_of(x):
    """(1 NT) -> (1, 'NT'); nil -> None"""
    x = unquote(x)
    if x is None:
        return None
    return (int(x[0]), str(x[1]).upper())


def truth(x):
    return x is not None and x is not False
