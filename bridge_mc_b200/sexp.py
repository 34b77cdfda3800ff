"""Minimal Common-Lisp S-expression reader/printer for the host side.

The reference talks in S-expressions at every boundary of the hot path:
range defs arrive from the frontend as text and go through `read-from-string`
(rest.cl:279-285), bidding rules are quoted forms (bidding.cl:76-106), and
`mc-case` rows are `prin1`-ed (bridge.cl:3201-3209).  This module reads that
syntax into Python values:

    symbol  -> Sym (upper-cased str subclass; NIL -> None, T -> True)
    integer -> int,  ratio a/b -> fractions.Fraction
    string  -> LStr (str subclass, keeps case)
    list    -> list   ('x -> [QUOTE x], #'f -> [FUNCTION f], `x -> [BACKQUOTE x],
                       ,x -> [UNQUOTE x], ,@x -> [UNQUOTE-SPLICE x])
    dotted pair (a . b) -> Dotted(list_of_heads, tail)
"""
from __future__ import annotations

from fractions import Fraction
import re


class Sym(str):
    """A Lisp symbol (stored upper-case, like the default readtable)."""

    def __repr__(self):
        return f"Sym({str.__repr__(self)})"


class LStr(str):
    """A Lisp string literal."""

    def __repr__(self):
        return f"LStr({str.__repr__(self)})"


class Dotted(list):
    """(a b . c): list of heads with `.tail`."""

    def __init__(self, heads, tail):
        super().__init__(heads)
        self.tail = tail

    def __repr__(self):
        return f"Dotted({list.__repr__(self)}, {self.tail!r})"


_INT = re.compile(r"^[+-]?\d+\.?$")
_RATIO = re.compile(r"^[+-]?\d+/\d+$")
_DELIMS = set("()'`,\";")


class SexpError(ValueError):
    pass


def _atom(tok: str):
    if _INT.match(tok):
        return int(tok.rstrip("."))
    if _RATIO.match(tok):
        n, d = tok.split("/")
        return Fraction(int(n), int(d))
    up = tok.upper()
    if up == "NIL":
        return None
    if up == "T":
        return True
    return Sym(up)


class _Reader:
    def __init__(self, text: str):
        self.s = text
        self.i = 0

    def _skip(self):
        s, n = self.s, len(self.s)
        while self.i < n:
            c = s[self.i]
            if c.isspace():
                self.i += 1
            elif c == ";":
                while self.i < n and s[self.i] != "\n":
                    self.i += 1
            elif s.startswith("#|", self.i):
                j = s.find("|#", self.i + 2)
                self.i = n if j < 0 else j + 2
            else:
                break

    def eof(self) -> bool:
        self._skip()
        return self.i >= len(self.s)

    def read(self):
        self._skip()
        s = self.s
        if self.i >= len(s):
            raise SexpError("unexpected end of input")
        c = s[self.i]
        if c == "(":
            self.i += 1
            items = []
            while True:
                self._skip()
                if self.i >= len(s):
                    raise SexpError("unbalanced '('")
                if s[self.i] == ")":
                    self.i += 1
                    return items
                if s[self.i] == "." and self.i + 1 < len(s) and (s[self.i + 1].isspace() or s[self.i + 1] == "("):
                    self.i += 1
                    tail = self.read()
                    self._skip()
                    if self.i >= len(s) or s[self.i] != ")":
                        raise SexpError("bad dotted pair")
                    self.i += 1
                    if isinstance(tail, list) and not isinstance(tail, Dotted):
                        return items + tail          # (a . (b c)) == (a b c)
                    if tail is None:
                        return items
                    return Dotted(items, tail)
                items.append(self.read())
        if c == ")":
            raise SexpError("unexpected ')'")
        if c == "'":
            self.i += 1
            return [Sym("QUOTE"), self.read()]
        if c == "`":
            self.i += 1
            return [Sym("BACKQUOTE"), self.read()]
        if c == ",":
            self.i += 1
            if self.i < len(s) and s[self.i] == "@":
                self.i += 1
                return [Sym("UNQUOTE-SPLICE"), self.read()]
            return [Sym("UNQUOTE"), self.read()]
        if c == '"':
            self.i += 1
            out = []
            while True:
                if self.i >= len(s):
                    raise SexpError("unterminated string")
                ch = s[self.i]
                if ch == "\\":
                    out.append(s[self.i + 1])
                    self.i += 2
                elif ch == '"':
                    self.i += 1
                    return LStr("".join(out))
                else:
                    out.append(ch)
                    self.i += 1
        if c == "#":
            nxt = s[self.i + 1] if self.i + 1 < len(s) else ""
            if nxt == "'":
                self.i += 2
                return [Sym("FUNCTION"), self.read()]
            if nxt == "\\":
                # character literal #\x or #\space
                j = self.i + 3
                while j < len(s) and not s[j].isspace() and s[j] not in "()":
                    j += 1
                name = s[self.i + 2:j]
                self.i = j
                return LStr({"space": " ", "newline": "\n"}.get(name.lower(), name))
        j = self.i
        while j < len(s) and not s[j].isspace() and s[j] not in _DELIMS:
            j += 1
        if c == "|":           # |sym| keeps case (rest.cl uses :|defs|)
            j = s.index("|", self.i + 1) + 1
        tok = s[self.i:j]
        if not tok:
            raise SexpError(f"cannot read at offset {self.i}: {s[self.i:self.i + 20]!r}")
        self.i = j
        return _atom(tok)


def read(text: str):
    """`read-from-string`: the first form in `text`."""
    return _Reader(text).read()


def read_all(text: str):
    """All top-level forms of a source text."""
    r = _Reader(text)
    out = []
    while not r.eof():
        out.append(r.read())
    return out


def read_all_with_lines(text: str):
    """[(line_number, form)] for every top-level form."""
    r = _Reader(text)
    out = []
    while not r.eof():
        line = text.count("\n", 0, r.i) + 1
        out.append((line, r.read()))
    return out


def dumps(x) -> str:
    """`prin1`-style printer (inverse of read for the value types above)."""
    if x is None:
        return "NIL"
    if x is True:
        return "T"
    if x is False:
        return "NIL"
    if isinstance(x, LStr):
        return '"' + x.replace("\\", "\\\\").replace('"', '\\"') + '"'
    if isinstance(x, Sym):
        return str(x)
    if isinstance(x, Fraction):
        return f"{x.numerator}/{x.denominator}" if x.denominator != 1 else str(x.numerator)
    if isinstance(x, int):
        return str(x)
    if isinstance(x, Dotted):
        return "(" + " ".join(dumps(e) for e in x) + " . " + dumps(x.tail) + ")"
    if isinstance(x, (list, tuple)):
        if len(x) == 2 and x[0] == "QUOTE" and isinstance(x[0], Sym):
            return "'" + dumps(x[1])
        return "(" + " ".join(dumps(e) for e in x) + ")"
    if isinstance(x, str):
        return str(x)
    raise TypeError(f"cannot print {type(x)}")
