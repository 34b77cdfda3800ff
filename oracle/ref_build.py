"""Reference-algorithm `build` with any number of ranged seats (bridge.cl:1292-1319), for distribution
checks of the GPU's reference-order mode.  TEST INFRASTRUCTURE ONLY.  The distgen enumeration and
sampling are the C restatement (bmc_oracle.c); the per-seat narrowing is the host mirror of infer-*
(bridge_mc_b200/infer.py), itself pinned by the reference's 14 KATs."""
from __future__ import annotations

import ctypes as C

from bridge_mc_b200.infer import BadRangeError, infer_distgen_params, supply
from . import oracle as O


def _popc(x):
    return bin(x).count("1")


def _cards(mask52):
    return [c for c in range(52) if (mask52 >> c) & 1]


_ROOT_CACHE = {}


def ref_build(consts, defs, state: "C.c_uint64"):
    """consts: 52-bit id-order masks; defs: list of dicts.  Returns the hands (52-bit masks) in `build`'s
    order: consts, sampled seats, deal-all seats -- or None where the reference signals an error."""
    cards = (1 << 52) - 1
    for m in consts:
        cards &= ~m
    hands = []
    rest = cards
    remaining = list(defs)
    first = True
    while remaining and (first or _popc(rest) > 13):
        first = False
        d, others = remaining[0], remaining[1:]
        try:
            params = infer_distgen_params(supply(_cards(rest)), d, others)
        except BadRangeError:
            return None
        kw = {k: (tuple(tuple(x) if isinstance(x, list) else x for x in v) if isinstance(v, list) else v) for k, v in params.items()}
        key = (rest, tuple(sorted(kw.items())))
        if len(remaining) == len(defs):          # the root generator is cached, like the reference's `rootgen` slot
            g = _ROOT_CACHE.get(key) or _ROOT_CACHE.setdefault(key, O.Distgen(rest, **kw))
        else:
            g = O.Distgen(rest, **kw)
        if g.total == 0:
            return None                       # (random 0) is an error in the reference
        h = g.rand_hand(state)
        hands.append(h)
        rest &= ~h
        remaining = others
    out = (C.c_uint64 * 4)()
    n = O.lib().orc_deal_all(rest, C.byref(state), out) if rest else 0
    return list(consts) + hands + [out[i] for i in range(n)]
