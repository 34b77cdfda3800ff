"""Pure-Python restatement of good-opening? (bidding.cl:60-74): substitute the 13 hand variables
into a rule form and evaluate it with Lisp semantics.  TEST INFRASTRUCTURE ONLY (small cases):
an evaluator independent of the C++ rule compiler in bridge_mc_b200/csrc/rules.hpp.

Values: nil = None, anything else is true; numbers are ints.  Supported forms are the ones the
reference's rule tables use: and or not cond if, < <= > >= = /=, find-if over (curry #'op n) and
lens / power / (subseq lens a b), quote, + - *.
"""
from __future__ import annotations


class IllegalCall(Exception):
    """evaluation reached a form whose head is not a function (bidding.cl:249 quirk)"""


def evaluate(form, feats):
    """feats: dict with lens[4], power[4], hcp, losers, balanced(bool).  Returns a Lisp value."""
    env = {"C": feats["lens"][0], "D": feats["lens"][1], "H": feats["lens"][2], "S": feats["lens"][3],
           "CPOWER": feats["power"][0], "DPOWER": feats["power"][1], "HPOWER": feats["power"][2], "SPOWER": feats["power"][3],
           "HCP": feats["hcp"], "LOOSERS": feats["losers"], "LOSERS": feats["losers"],
           "BALANCED": True if feats["balanced"] else None, "LENS": list(feats["lens"]), "POWER": list(feats["power"])}
    return _ev(form, env)


def _truth(v):
    return v is not None and v is not False


_CMP = {"<": lambda a, b: a < b, "<=": lambda a, b: a <= b, ">": lambda a, b: a > b, ">=": lambda a, b: a >= b,
        "=": lambda a, b: a == b, "/=": lambda a, b: a != b}


def _ev(x, env):
    if x is None:
        return None
    if x is True:
        return True
    if isinstance(x, int):
        return x
    if isinstance(x, str):                       # symbol
        if x in env:
            return env[x]
        raise KeyError(f"unbound variable {x}")
    if not isinstance(x, list) or not x:
        return None
    head = x[0]
    if not isinstance(head, str):
        raise IllegalCall(repr(x))
    args = x[1:]
    if head == "AND":
        v = True
        for a in args:
            v = _ev(a, env)
            if not _truth(v):
                return None
        return v
    if head == "OR":
        for a in args:
            v = _ev(a, env)
            if _truth(v):
                return v
        return None
    if head in ("NOT", "NULL"):
        return None if _truth(_ev(args[0], env)) else True
    if head == "COND":
        for clause in args:
            t = _ev(clause[0], env)
            if _truth(t):
                v = t
                for body in clause[1:]:
                    v = _ev(body, env)
                return v
        return None
    if head == "IF":
        if _truth(_ev(args[0], env)):
            return _ev(args[1], env)
        return _ev(args[2], env) if len(args) > 2 else None
    if head in _CMP:
        vals = [_ev(a, env) for a in args]
        ok = all(_CMP[head](a, b) for a, b in zip(vals, vals[1:]))
        return True if ok else None
    if head in ("+", "-", "*"):
        vals = [_ev(a, env) for a in args]
        if head == "+":
            return sum(vals)
        if head == "*":
            out = 1
            for v in vals:
                out *= v
            return out
        return -vals[0] if len(vals) == 1 else vals[0] - sum(vals[1:])
    if head == "QUOTE":
        return args[0]
    if head == "SUBSEQ":
        seq = _ev(args[0], env)
        return seq[_ev(args[1], env):(_ev(args[2], env) if len(args) > 2 else None)]
    if head == "FIND-IF":
        fn = args[0]                                 # (curry #'op n)
        assert fn[0] == "CURRY" and fn[1][0] == "FUNCTION", fn
        op, n = _CMP[fn[1][1]], _ev(fn[2], env)
        for e in _ev(args[1], env):
            if op(n, e):
                return e
        return None
    raise NotImplementedError(head)


def choose_bid(table, feats):
    """bidding.cl:146-150 over [(bid, form), ...]: first bid whose form is true, else None.
    Raises IllegalCall where the reference would signal an error before finding one."""
    for bid, form in table:
        if _truth(evaluate(form, feats)):
            return bid
    return None
