"""Host mirror of compscore.cl (duplicate scoring) on top of bmc_score /
bmc_match_points.  String parsing stays on the host, as the parser stays in Lisp
in the reference; every score is computed on the GPU (no CPU arithmetic path).
"""
from __future__ import annotations

import numpy as np

from ._lib import Contract as CContract
from .engine import Engine

_SUIT_OPTS = ("C", "D", "H", "S", "NT", "♣", "♦", "♥", "♠")
_SUIT_CODE = {"C": 0, "♣": 0, "D": 1, "♦": 1, "H": 2, "♥": 2, "S": 3, "♠": 3, "NT": 4, None: 4}
_SUIT_UNICODE = {"C": "♣", "D": "♦", "H": "♥", "S": "♠", "NT": "NT", "BA": "BA"}
_DECL = {"N": 0, "E": 1, "S": 2, "W": 3, None: -1}
_DBL = {None: 0, "X": 1, "XX": 2}


def scanstr(opts, text):
    """compscore.cl:18-24: (first option that prefixes `text` case-insensitively, rest) or (None, text)."""
    for opt in opts:
        o = str(opt)
        if len(text) >= len(o) and text[:len(o)].lower() == o.lower():
            return opt, text[len(o):]
    return None, text


class Contract:
    """compscore.cl:38-65: "<level><suit>[declarer][x|xx][=|+n|-n]", e.g. "4HSx+1"."""

    def __init__(self, text: str):
        self.level, rest = scanstr((1, 2, 3, 4, 5, 6, 7), text)
        self.suit, rest = scanstr(_SUIT_OPTS, rest)
        self.declarer, rest = scanstr(("N", "E", "S", "W"), rest)
        self.double, rest = scanstr(("XX", "X"), rest)
        state, after = scanstr(("=", "+", "-"), rest)
        if state and state != "=":
            self.outcome = f"{state}{int(after)}"
        else:
            self.outcome = "="

    # compscore.cl:67-75
    def suit_unicode(self):
        return _SUIT_UNICODE.get(self.suit)

    def __str__(self):
        su = self.suit_unicode()
        return f"{self.level}{su if su is not None else 'NIL'}{self.declarer or ''}" \
               f"{(self.double or '').lower()}{self.outcome if self.outcome is not None else ''}"

    __repr__ = __str__

    def with_score(self, tricks: int):
        """compscore.cl:77-81: set the outcome from a trick count."""
        n = tricks - (6 + self.level)
        self.outcome = "=" if n == 0 else n
        return self

    def suit_sym(self):
        """compscore.cl:83-84"""
        return None if self.suit in ("NT", "BA") else self.suit

    def tricks(self) -> int:
        o = self.outcome
        return self.level + 6 + (0 if o == "=" else int(o))

    def c_struct(self) -> CContract:
        if self.level is None:
            raise ValueError("contract has no level")
        return CContract(self.level, _SUIT_CODE[self.suit], _DECL[self.declarer], _DBL[self.double])


def contract_score(suit, tricks, vulnerable=False, double=None, level=None, engine: Engine | None = None) -> int:
    """bridge.cl:3045-3069 contract-score (suit tricks &key vulnerable double level)."""
    eng = engine or Engine.default()
    code = _SUIT_CODE[suit.upper() if isinstance(suit, str) else suit]
    dbl = _DBL[double.upper() if isinstance(double, str) else double]
    c = CContract(level or 0, code, -1, dbl)
    return int(eng.score([c], [1 if vulnerable else 0], np.array([[tricks]], dtype=np.uint8))[0, 0])


def base_score(contract: Contract, vulnerable=False, engine: Engine | None = None) -> int:
    """compscore.cl:94-100: the contract's score, negated for an E/W declarer."""
    eng = engine or Engine.default()
    t = contract.tricks()
    if not 0 <= t <= 13:
        raise ValueError(f"{contract}: {t} tricks")
    return int(eng.score([contract.c_struct()], [1 if vulnerable else 0], np.array([[t]], dtype=np.uint8))[0, 0])


def match_points(a: Contract, b: Contract, vulnerable=None, engine: Engine | None = None) -> int:
    """compscore.cl:127-131: IMPs of a against b; vulnerable in (None, 'ns', 'ew', 'both')."""
    eng = engine or Engine.default()
    v = (vulnerable or "").lower() if isinstance(vulnerable, str) else vulnerable
    sa = base_score(a, vulnerable=v in ("ns", "both"), engine=eng)
    sb = base_score(b, vulnerable=v in ("ew", "both"), engine=eng)
    imps, status = eng.match_points(np.array([sa], np.int32), np.array([sb], np.int32))
    if status[0]:
        raise ValueError(f"score difference {sa - sb} is off the IMP table")   # the reference errors on (* nil ..)
    return int(imps[0])
