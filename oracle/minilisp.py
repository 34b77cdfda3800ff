"""TEST INFRASTRUCTURE -- a small Common Lisp evaluator, enough to run the reference's OWN source files
(/root/reference/backend/*.cl) in place, in the build container.  There is no Lisp in the image
(sbcl / ecl / clisp ... are all absent), so this is how the reference itself is executed here: its inline
`(test ...)` forms are replayed, and golden input/output vectors are generated from its functions
(oracle/run_reference.py).  It reads the reference where it lies; no reference source is copied.  Nothing
under bridge_mc_b200/ may import this file, and nothing on the GPU box needs it (the vectors it produces are
committed under tests/golden/).

Scope: the subset of CL the reference uses -- special forms (quote function if cond and or let let* flet labels
lambda setf progn loop destructuring-bind with-slots eval backquote), defun / defmacro with &optional &rest &key
&body, single-inheritance-free CLOS (defclass with :initarg :initform :reader :accessor; defmethod dispatching on
the classes of the required arguments; a user `initialize-instance` method REPLACES the default one, as in CLOS,
because the reference's never call-next-method), hash tables with `equal` keys, and ~90 list / number / string
functions.  Lisp lists are Python lists (nil = None; an empty Python list never escapes), so `cdr` copies:
destructive operations are supported where the reference uses them -- (setf (nth ..)), (setf (cdr (last x)) ..),
(setf (car/first..fourth ..)), pop / push on any place -- and `sort` returns a fresh list.

Validation of the evaluator itself: every inline test of bridge.cl that the reference's authors keep green must
pass here (oracle/run_reference.py prints the tally).
"""
from __future__ import annotations

import random as _random
import sys
from fractions import Fraction

sys.setrecursionlimit(4000000)


class LispError(Exception):
    pass


class Sym(str):
    __slots__ = ()

    def __repr__(self):
        return str(self)


class LStr(str):
    """a Lisp string: identity matters for `eq` (hand ids), so never interned"""
    __slots__ = ()


class Char(str):
    __slots__ = ()


class Dotted(list):
    """(a b . c) with an atom c: the heads, plus `.tail`"""

    def __init__(self, heads, tail):
        super().__init__(heads)
        self.tail = tail


_SYMS: dict = {}


def S(name):
    s = _SYMS.get(name)
    if s is None:
        s = _SYMS[name] = Sym(name)
    return s


QUOTE, FUNCTION, BACKQUOTE, UNQUOTE, SPLICE = S("QUOTE"), S("FUNCTION"), S("BACKQUOTE"), S("UNQUOTE"), S("UNQUOTE-SPLICE")
LAMBDA, T_ = S("LAMBDA"), True
AMP = {S("&OPTIONAL"), S("&REST"), S("&KEY"), S("&BODY"), S("&ALLOW-OTHER-KEYS")}


# ---------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------
class Reader:
    DELIMS = set("()'`,\";")

    def __init__(self, text):
        self.s, self.i = text, 0

    def skip(self):
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

    def eof(self):
        self.skip()
        return self.i >= len(self.s)

    def read(self):
        self.skip()
        s = self.s
        if self.i >= len(s):
            raise LispError("eof")
        c = s[self.i]
        if c == "(":
            self.i += 1
            items = []
            while True:
                self.skip()
                if self.i >= len(s):
                    raise LispError("unbalanced (")
                if s[self.i] == ")":
                    self.i += 1
                    return items if items else None
                if s[self.i] == "." and self.i + 1 < len(s) and (s[self.i + 1].isspace() or s[self.i + 1] in "(,'`"):
                    self.i += 1
                    tail = self.read()
                    self.skip()
                    if s[self.i] != ")":
                        raise LispError("bad dotted list")
                    self.i += 1
                    if tail is None:
                        return items
                    if isinstance(tail, list) and not isinstance(tail, Dotted):
                        return items + tail
                    return Dotted(items, tail)
                items.append(self.read())
        if c == ")":
            raise LispError("unexpected )")
        if c == "'":
            self.i += 1
            return [QUOTE, self.read()]
        if c == "`":
            self.i += 1
            return [BACKQUOTE, self.read()]
        if c == ",":
            self.i += 1
            if s[self.i] == "@":
                self.i += 1
                return [SPLICE, self.read()]
            return [UNQUOTE, self.read()]
        if c == '"':
            self.i += 1
            out = []
            while True:
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
            nxt = s[self.i + 1]
            if nxt == "'":
                self.i += 2
                return [FUNCTION, self.read()]
            if nxt == "\\":
                j = self.i + 3
                while j < len(s) and not s[j].isspace() and s[j] not in "()":
                    j += 1
                name = s[self.i + 2:j]
                self.i = j
                return Char({"space": " ", "newline": "\n", "tab": "\t"}.get(name.lower(), name))
        j = self.i
        if c == "|":
            j = s.index("|", self.i + 1) + 1
        else:
            while j < len(s) and not s[j].isspace() and s[j] not in self.DELIMS:
                j += 1
        tok = s[self.i:j]
        self.i = j
        return atom(tok)


def atom(tok):
    t = tok
    try:
        return int(t.rstrip(".")) if t.rstrip(".").lstrip("+-").isdigit() else _not_int(t)
    except ValueError:
        return _not_int(t)


def _not_int(t):
    if "/" in t:
        a, _, b = t.partition("/")
        if a.lstrip("+-").isdigit() and b.isdigit():
            return Fraction(int(a), int(b))
    up = t.upper()
    if up == "NIL":
        return None
    if up == "T":
        return True
    if t.startswith("|") or t.startswith(":|"):
        return S(t.replace("|", ""))
    return S(up)


def read_all_with_lines(text):
    r = Reader(text)
    out = []
    while not r.eof():
        line = text.count("\n", 0, r.i) + 1
        out.append((line, r.read()))
    return out


def read_from_string(text):
    return Reader(text).read()


# ---------------------------------------------------------------------------
# printer (prin1-ish; enough for diagnostics and `format`)
# ---------------------------------------------------------------------------
def lisp_repr(x, readably=True):
    if x is None:
        return "NIL"
    if x is True:
        return "T"
    if isinstance(x, Dotted):
        return "(" + " ".join(lisp_repr(e, readably) for e in x) + " . " + lisp_repr(x.tail, readably) + ")"
    if isinstance(x, list):
        return "(" + " ".join(lisp_repr(e, readably) for e in x) + ")"
    if isinstance(x, Sym):
        return str(x)
    if isinstance(x, Char):
        return ("#\\" + x) if readably else str(x)
    if isinstance(x, str):
        return '"' + x + '"' if readably else str(x)
    if isinstance(x, Instance):
        it = x.cls.interp
        po = it.generics.get(S("PRINT-OBJECT")) if it is not None else None
        if po is not None and po.has_method_for(x.cls.name):
            stream = StringStream()
            po.fn(x, stream)
            return "".join(stream.parts)
        return "#<%s>" % x.cls.name
    return str(x)


class StringStream:
    def __init__(self):
        self.parts = []


def lisp_format(ctrl, args):
    out, i, args = [], 0, list(args)
    while i < len(ctrl):
        c = ctrl[i]
        if c != "~":
            out.append(c)
            i += 1
            continue
        i += 1
        j = i
        while ctrl[j] in "0123456789,":
            j += 1
        d = ctrl[j].upper()
        i = j + 1
        if d == "A":
            out.append(lisp_repr(args.pop(0), False))
        elif d == "S":
            out.append(lisp_repr(args.pop(0), True))
        elif d == "X":
            out.append(format(args.pop(0), "x").upper())
        elif d == "D":
            out.append(str(args.pop(0)))
        elif d == "%":
            out.append("\n")
        elif d == "F":
            v = args.pop(0)
            out.append("%.2f" % float(v))
        elif d == "{":
            k = ctrl.index("~}", i)
            inner = ctrl[i:k]
            lst = list(args.pop(0) or [])
            while lst:
                n_before = len(lst)
                out.append(lisp_format_consume(inner, lst))
                if len(lst) == n_before:
                    break
            i = k + 2
        else:
            raise LispError("format directive ~" + d)
    return LStr("".join(out))


def lisp_format_consume(ctrl, lst):
    n = ctrl.upper().count("~A") + ctrl.upper().count("~S")
    args = [lst.pop(0) for _ in range(min(n, len(lst)))]
    return lisp_format(ctrl, args)


# ---------------------------------------------------------------------------
# objects
# ---------------------------------------------------------------------------
class LClass:
    def __init__(self, name, supers, slots, interp=None):
        self.name, self.supers, self.slots = name, supers, slots     # slots: [(name, initargs, initform, has_initform)]
        self.interp = interp


class Instance:
    __slots__ = ("cls", "slots")

    def __init__(self, cls):
        self.cls, self.slots = cls, {}


class SlotRef:
    __slots__ = ("obj", "name")

    def __init__(self, obj, name):
        self.obj, self.name = obj, name


class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None, vars=None):
        self.vars = vars if vars is not None else {}
        self.parent = parent

    def find(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e
            e = e.parent
        return None


class FEnv:
    __slots__ = ("fns", "parent")

    def __init__(self, parent=None):
        self.fns, self.parent = {}, parent

    def find(self, name):
        e = self
        while e is not None:
            if name in e.fns:
                return e.fns[name]
            e = e.parent
        return None


def Closure(interp, params, body, env, fenv, name="LAMBDA"):
    """A Lisp function as a plain Python function: CPython runs Python-to-Python calls without C recursion, which
    the reference's deeply recursive list helpers (`filter`, `fold` over lists of thousands of tricks) need."""
    plan = parse_lambda_list(params)

    def closure(*args):
        new = Env(env)
        interp.bind_args(plan, list(args), new, fenv_cell[0], name)
        out = None
        for b in body:
            out = interp.eval(b, new, fenv_cell[0])
        return out
    fenv_cell = [fenv]
    closure.fenv_cell = fenv_cell
    closure.lisp_name = name
    return closure


def parse_lambda_list(params):
    req, opt, rest, keys, mode = [], [], None, [], "req"
    for p in params or []:
        if isinstance(p, Sym) and p in AMP:
            mode = {"&OPTIONAL": "opt", "&REST": "rest", "&BODY": "rest", "&KEY": "key", "&ALLOW-OTHER-KEYS": "aok"}[str(p)]
            continue
        if mode == "req":
            req.append(p)
        elif mode == "opt":
            opt.append((p[0], p[1] if len(p) > 1 else None) if isinstance(p, list) else (p, None))
        elif mode == "rest":
            rest = p
        elif mode == "key":
            keys.append((p[0], p[1] if len(p) > 1 else None) if isinstance(p, list) else (p, None))
    return req, opt, rest, keys


def is_kw(x):
    return isinstance(x, Sym) and x.startswith(":")


def truthy(x):
    return x is not None and x is not False


def lst(x):
    """Python list -> Lisp list (empty -> nil)"""
    return x if x else None


def eq(a, b):
    if a is b:
        return True
    ta, tb = type(a), type(b)
    if ta is tb and ta in (Sym, int, Char, Fraction):
        return a == b
    if a is None and b is None:
        return True
    return False


def equal(a, b):
    if eq(a, b):
        return True
    if isinstance(a, list) and isinstance(b, list):
        if isinstance(a, Dotted) != isinstance(b, Dotted) or (isinstance(a, Dotted) and not equal(a.tail, b.tail)):
            return False
        return len(a) == len(b) and all(equal(x, y) for x, y in zip(a, b))
    if isinstance(a, str) and isinstance(b, str) and not isinstance(a, Sym) and not isinstance(b, Sym):
        return str(a) == str(b) and isinstance(a, Char) == isinstance(b, Char)
    if isinstance(a, (int, Fraction)) and isinstance(b, (int, Fraction)) and type(a) is type(b):
        return a == b
    return False


def hash_key(x):
    if isinstance(x, list):
        return tuple(hash_key(e) for e in x)
    if isinstance(x, Instance):
        return ("#i", id(x))
    if callable(x) and not isinstance(x, str):
        return ("#f", id(x))
    if isinstance(x, Sym):
        return ("#s", str(x))
    if isinstance(x, str):
        return ("#str", str(x))
    return x


class HashTable:
    def __init__(self):
        self.d = {}
        self.keep = []


# ---------------------------------------------------------------------------
# the evaluator
# ---------------------------------------------------------------------------
class Interp:
    def __init__(self, seed=1):
        self.gfun, self.macros, self.gvar, self.classes, self.generics = {}, {}, {}, {}, {}
        self.rng = _random.Random(seed)
        self.out = []
        self.expansions = {}
        self._keep = []
        self.special = {
            "QUOTE": self.sf_quote, "FUNCTION": self.sf_function, "IF": self.sf_if, "COND": self.sf_cond,
            "AND": self.sf_and, "OR": self.sf_or, "LET": self.sf_let, "LET*": self.sf_letstar, "FLET": self.sf_flet,
            "LABELS": self.sf_labels, "LAMBDA": self.sf_lambda, "SETF": self.sf_setf, "SETQ": self.sf_setf,
            "PROGN": self.sf_progn, "LOOP": self.sf_loop, "DESTRUCTURING-BIND": self.sf_dbind, "WITH-SLOTS": self.sf_with_slots,
            "BACKQUOTE": self.sf_backquote, "DEFUN": self.sf_defun, "DEFMACRO": self.sf_defmacro, "DEFCLASS": self.sf_defclass,
            "DEFMETHOD": self.sf_defmethod, "DEFGENERIC": self.sf_defgeneric, "DEFVAR": self.sf_defvar,
            "DEFPARAMETER": self.sf_defvar, "DEFINE-CONDITION": self.sf_defcondition, "DECLARE": lambda f, e, fe: None,
            "POP": self.sf_pop, "PUSH": self.sf_push, "WHEN": self.sf_when, "UNLESS": self.sf_unless, "INCF": self.sf_incf,
            "MULTIPLE-VALUE-BIND": self.sf_mvb, "IGNORE-ERRORS": self.sf_ignore_errors,
        }
        install_builtins(self)

    # --- function lookup --------------------------------------------------
    def fn(self, name, fenv=None):
        if fenv is not None:
            f = fenv.find(name)
            if f is not None:
                return f
        f = self.gfun.get(name)
        if f is not None:
            return f
        g = self.generics.get(name)
        if g is not None:
            return g.fn
        raise LispError("undefined function " + str(name))

    def bind_args(self, plan, args, env, fenv, name):
        req, opt, rest, keys = plan
        n = len(req)
        if len(args) < n:
            raise LispError("%s: too few arguments (%d < %d)" % (name, len(args), n))
        v = env.vars
        for p, a in zip(req, args):
            v[p] = a
        i = n
        for p, default in opt:
            if i < len(args):
                v[p] = args[i]
                i += 1
            else:
                v[p] = self.eval(default, env, fenv) if default is not None else None
        restargs = args[i:]
        if rest is not None:
            v[rest] = lst(list(restargs))
        elif not keys and restargs:
            raise LispError("%s: too many arguments" % name)
        if keys:
            given = {}
            for k in range(0, len(restargs) - 1, 2):
                if restargs[k] not in given:
                    given[restargs[k]] = restargs[k + 1]
            for p, default in keys:
                kw = S(":" + str(p))
                if kw in given:
                    v[p] = given[kw]
                else:
                    v[p] = self.eval(default, env, fenv) if default is not None else None

    # --- eval -------------------------------------------------------------
    def eval(self, f, env, fenv=None):
        if isinstance(f, Sym):
            if f.startswith(":"):
                return f
            e = env.find(f) if env is not None else None
            if e is not None:
                val = e.vars[f]
                if type(val) is SlotRef:
                    try:
                        return val.obj.slots[val.name]
                    except KeyError:
                        raise LispError("unbound slot %s" % val.name)
                return val
            if f in self.gvar:
                return self.gvar[f]
            raise LispError("unbound variable " + str(f))
        if not isinstance(f, list):
            return f
        head = f[0]
        if isinstance(head, Sym):
            sf = self.special.get(head)
            if sf is not None and (fenv is None or fenv.find(head) is None):
                return sf(f, env, fenv)
            m = self.macros.get(head)
            if m is not None and (fenv is None or fenv.find(head) is None):
                key = id(f)
                exp = self.expansions.get(key)
                if exp is None or exp[0] is not f:
                    exp = (f, m(*f[1:]))
                    self.expansions[key] = exp
                return self.eval(exp[1], env, fenv)
            fn = self.fn(head, fenv)
        elif isinstance(head, list) and head and head[0] is LAMBDA:
            fn = self.eval(head, env, fenv)
        else:
            raise LispError("bad function position: " + lisp_repr(head))
        args = [self.eval(a, env, fenv) for a in f[1:]]
        return fn(*args)

    def progn(self, body, env, fenv):
        out = None
        for b in body:
            out = self.eval(b, env, fenv)
        return out

    # --- special forms ------------------------------------------------------
    def sf_quote(self, f, env, fenv):
        return f[1]

    def sf_function(self, f, env, fenv):
        x = f[1]
        if isinstance(x, list):
            return self.eval(x, env, fenv)
        return self.fn(x, fenv)

    def sf_if(self, f, env, fenv):
        if truthy(self.eval(f[1], env, fenv)):
            return self.eval(f[2], env, fenv)
        return self.eval(f[3], env, fenv) if len(f) > 3 else None

    def sf_when(self, f, env, fenv):
        return self.progn(f[2:], env, fenv) if truthy(self.eval(f[1], env, fenv)) else None

    def sf_unless(self, f, env, fenv):
        return self.progn(f[2:], env, fenv) if not truthy(self.eval(f[1], env, fenv)) else None

    def sf_cond(self, f, env, fenv):
        for clause in f[1:]:
            v = self.eval(clause[0], env, fenv)
            if truthy(v):
                return self.progn(clause[1:], env, fenv) if len(clause) > 1 else v
        return None

    def sf_and(self, f, env, fenv):
        v = True
        for x in f[1:]:
            v = self.eval(x, env, fenv)
            if not truthy(v):
                return None
        return v

    def sf_or(self, f, env, fenv):
        for x in f[1:]:
            v = self.eval(x, env, fenv)
            if truthy(v):
                return v
        return None

    def sf_let(self, f, env, fenv):
        new = Env(env)
        for b in f[1] or []:
            if isinstance(b, list):
                new.vars[b[0]] = self.eval(b[1], env, fenv) if len(b) > 1 else None
            else:
                new.vars[b] = None
        return self.progn(f[2:], new, fenv)

    def sf_letstar(self, f, env, fenv):
        new = Env(env)
        for b in f[1] or []:
            if isinstance(b, list):
                new.vars[b[0]] = self.eval(b[1], new, fenv) if len(b) > 1 else None
            else:
                new.vars[b] = None
        return self.progn(f[2:], new, fenv)

    def sf_flet(self, f, env, fenv):
        new = FEnv(fenv)
        for name, params, *body in f[1]:
            new.fns[name] = Closure(self, params, body, env, fenv, str(name))
        return self.progn(f[2:], env, new)

    def sf_labels(self, f, env, fenv):
        new = FEnv(fenv)
        for name, params, *body in f[1]:
            new.fns[name] = Closure(self, params, body, env, new, str(name))
        return self.progn(f[2:], env, new)

    def sf_lambda(self, f, env, fenv):
        return Closure(self, f[1], f[2:], env, fenv)

    def sf_progn(self, f, env, fenv):
        return self.progn(f[1:], env, fenv)

    def sf_mvb(self, f, env, fenv):
        new = Env(env)
        v = self.eval(f[2], env, fenv)
        for i, name in enumerate(f[1]):
            new.vars[name] = v if i == 0 else None
        return self.progn(f[3:], new, fenv)

    def sf_ignore_errors(self, f, env, fenv):
        try:
            return self.progn(f[1:], env, fenv)
        except LispError:
            return None

    def sf_backquote(self, f, env, fenv):
        return self.quasi(f[1], env, fenv)

    def quasi(self, x, env, fenv):
        if isinstance(x, list):
            if len(x) == 2 and x[0] is UNQUOTE:
                return self.eval(x[1], env, fenv)
            if len(x) >= 3 and x[-2] is UNQUOTE and not isinstance(x, Dotted):      # (a b . ,form)
                heads = self.quasi(x[:-2], env, fenv) or []
                tail = self.eval(x[-1], env, fenv)
                if tail is None:
                    return lst(heads)
                if isinstance(tail, list) and not isinstance(tail, Dotted):
                    return heads + tail
                return Dotted(heads, tail)
            out = []
            for e in x:
                if isinstance(e, list) and len(e) == 2 and e[0] is SPLICE:
                    out.extend(self.eval(e[1], env, fenv) or [])
                else:
                    out.append(self.quasi(e, env, fenv))
            return lst(out)
        return x

    def sf_dbind(self, f, env, fenv):
        new = Env(env)
        val = self.eval(f[2], env, fenv)
        self.destructure(f[1], val, new, fenv)
        return self.progn(f[3:], new, fenv)

    def destructure(self, pattern, val, env, fenv):
        val = val or []
        i = 0
        mode = "req"
        for p in pattern or []:
            if isinstance(p, Sym) and p in AMP:
                mode = str(p)
                continue
            if mode == "&REST" or mode == "&BODY":
                env.vars[p] = lst(list(val[i:]))
            elif mode == "&OPTIONAL":
                name, default = (p[0], p[1] if len(p) > 1 else None) if isinstance(p, list) else (p, None)
                env.vars[name] = val[i] if i < len(val) else (self.eval(default, env, fenv) if default is not None else None)
                i += 1
            elif mode == "&KEY":
                name, default = (p[0], p[1] if len(p) > 1 else None) if isinstance(p, list) else (p, None)
                kw, found = S(":" + str(name)), False
                for k in range(i, len(val) - 1, 2):
                    if val[k] == kw and isinstance(val[k], Sym):
                        env.vars[name], found = val[k + 1], True
                        break
                if not found:
                    env.vars[name] = self.eval(default, env, fenv) if default is not None else None
            elif mode == "&ALLOW-OTHER-KEYS":
                pass
            else:
                if i >= len(val):
                    raise LispError("destructuring-bind: too few elements")
                if isinstance(p, list):
                    self.destructure(p, val[i], env, fenv)
                else:
                    env.vars[p] = val[i]
                i += 1
        if mode == "req" and i < len(val):
            raise LispError("destructuring-bind: too many elements")

    def sf_with_slots(self, f, env, fenv):
        obj = self.eval(f[2], env, fenv)
        if not isinstance(obj, Instance):
            raise LispError("with-slots on " + lisp_repr(obj))
        new = Env(env)
        for s in f[1] or []:
            if isinstance(s, list):
                new.vars[s[0]] = SlotRef(obj, s[1])
            else:
                new.vars[s] = SlotRef(obj, s)
        return self.progn(f[3:], new, fenv)

    # --- places -------------------------------------------------------------
    def sf_setf(self, f, env, fenv):
        val = None
        for i in range(1, len(f) - 1, 2):
            val = self.set_place(f[i], self.eval(f[i + 1], env, fenv), env, fenv)
        return val

    def set_place(self, place, val, env, fenv):
        if isinstance(place, Sym):
            e = env.find(place) if env is not None else None
            if e is not None:
                cur = e.vars[place]
                if type(cur) is SlotRef:
                    cur.obj.slots[cur.name] = val
                else:
                    e.vars[place] = val
            else:
                self.gvar[place] = val
            return val
        head = place[0]
        m = self.macros.get(head)
        if m is not None:
            return self.set_place(m(*place[1:]), val, env, fenv)
        h = str(head)
        if h == "SLOT-VALUE":
            obj, name = self.eval(place[1], env, fenv), self.eval(place[2], env, fenv)
            obj.slots[name] = val
        elif h == "GETHASH":
            key, table = self.eval(place[1], env, fenv), self.eval(place[2], env, fenv)
            table.d[hash_key(key)] = (key, val)
        elif h in ("NTH", "FIRST", "SECOND", "THIRD", "FOURTH", "CAR"):
            if h == "NTH":
                idx, target = self.eval(place[1], env, fenv), self.eval(place[2], env, fenv)
            else:
                idx, target = {"FIRST": 0, "CAR": 0, "SECOND": 1, "THIRD": 2, "FOURTH": 3}[h], self.eval(place[1], env, fenv)
            if not isinstance(target, list) or idx >= len(target):
                raise LispError("setf nth beyond the list")
            target[idx] = val
        elif h == "CDR":
            inner = place[1]
            if isinstance(inner, list) and inner and str(inner[0]) == "LAST" and len(inner) == 2:
                target = self.eval(inner[1], env, fenv)            # (setf (cdr (last x)) v): append in place
                if not isinstance(target, list):
                    raise LispError("setf cdr last of nil")
                target.extend(val or [])
            else:
                target = self.eval(inner, env, fenv)
                if not isinstance(target, list):
                    raise LispError("setf cdr of nil")
                target[1:] = val or []
        else:
            # a reader / accessor used as a place
            obj = self.eval(place[1], env, fenv)
            g = self.generics.get(head)
            if isinstance(obj, Instance) and g is not None and g.accessor_slot(obj) is not None:
                obj.slots[g.accessor_slot(obj)] = val
            else:
                raise LispError("setf: unsupported place " + lisp_repr(place))
        return val

    def sf_pop(self, f, env, fenv):
        cur = self.eval(f[1], env, fenv)
        if cur is None:
            return None
        self.set_place(f[1], lst(cur[1:]), env, fenv)
        return cur[0]

    def sf_push(self, f, env, fenv):
        item = self.eval(f[1], env, fenv)
        cur = self.eval(f[2], env, fenv)
        return self.set_place(f[2], [item] + (cur or []), env, fenv)

    def sf_incf(self, f, env, fenv):
        cur = self.eval(f[1], env, fenv)
        d = self.eval(f[2], env, fenv) if len(f) > 2 else 1
        return self.set_place(f[1], cur + d, env, fenv)

    # --- definitions ----------------------------------------------------------
    def sf_defun(self, f, env, fenv):
        body = f[3:]
        if body and isinstance(body[0], str) and not isinstance(body[0], Sym) and len(body) > 1:
            body = body[1:]
        self.gfun[f[1]] = Closure(self, f[2], body, None, None, str(f[1]))
        return f[1]

    def sf_defmacro(self, f, env, fenv):
        self.macros[f[1]] = Closure(self, f[2], f[3:], None, None, str(f[1]))
        self.expansions.clear()
        return f[1]

    def sf_defvar(self, f, env, fenv):
        self.gvar[f[1]] = self.eval(f[2], env, fenv) if len(f) > 2 else None
        return f[1]

    def sf_defclass(self, f, env, fenv):
        name, supers, slotdefs = f[1], f[2] or [], f[3] or []
        slots = []
        for sd in slotdefs:
            if not isinstance(sd, list):
                sd = [sd]
            sname, opts = sd[0], sd[1:]
            initargs, initform, has = [], None, False
            for k in range(0, len(opts) - 1, 2):
                o, v = str(opts[k]), opts[k + 1]
                if o == ":INITARG":
                    initargs.append(v)
                elif o == ":INITFORM":
                    initform, has = v, True
                elif o in (":READER", ":ACCESSOR"):
                    self.generic(v).add_reader(name, sname)
            slots.append((sname, initargs, initform, has))
        self.classes[name] = LClass(name, supers, slots, self)
        return name

    def sf_defcondition(self, f, env, fenv):
        return self.sf_defclass([f[0], f[1], None, f[3]], env, fenv)

    def sf_defgeneric(self, f, env, fenv):
        self.generic(f[1])
        return f[1]

    def generic(self, name):
        g = self.generics.get(name)
        if g is None:
            g = self.generics[name] = Generic(self, name)
        return g

    def sf_defmethod(self, f, env, fenv):
        name, params, body = f[1], f[2], f[3:]
        specs, plain = [], []
        seen_amp = False
        for p in params or []:
            if isinstance(p, Sym) and p in AMP:
                seen_amp = True
            if not seen_amp and isinstance(p, list):
                plain.append(p[0])
                specs.append(p[1])
            else:
                plain.append(p)
                if not seen_amp:
                    specs.append(None)
        self.generic(name).add_method(specs, Closure(self, plain, body, None, None, str(name)))
        return name

    # --- instances -------------------------------------------------------------
    def class_of(self, x):
        if isinstance(x, Instance):
            return x.cls.name
        if x is None or isinstance(x, list):
            return S("LIST")
        if isinstance(x, Sym) or x is True:
            return S("SYMBOL")
        if isinstance(x, (int, Fraction)):
            return S("NUMBER")
        if isinstance(x, str):
            return S("STRING")
        if callable(x):
            return S("FUNCTION")
        return S("T")

    def make_instance(self, cname, *initargs):
        cls = self.classes.get(cname)
        if cls is None:
            raise LispError("no class " + str(cname))
        obj = Instance(cls)
        init = self.generics.get(S("INITIALIZE-INSTANCE"))
        if init is not None and init.has_method_for(cname):
            init.fn(obj, *initargs)                    # a user primary method replaces the default (no call-next-method)
            return obj
        given = {}
        for k in range(0, len(initargs) - 1, 2):
            if initargs[k] not in given:
                given[initargs[k]] = initargs[k + 1]
        for sname, ias, initform, has in cls.slots:
            done = False
            for ia in ias:
                if ia in given:
                    obj.slots[sname] = given[ia]
                    done = True
                    break
            if not done and has:
                obj.slots[sname] = self.eval(initform, None, None)
        return obj

    # --- loop ----------------------------------------------------------------------
    LOOP_KW = {"FOR", "REPEAT", "WHILE", "UNTIL", "COLLECT", "APPEND", "SUM", "DO", "AS"}

    def sf_loop(self, f, env, fenv):
        toks = f[1:]
        iters, body_clauses = [], []
        i = 0

        def kw(x):
            return isinstance(x, Sym) and not x.startswith(":") and str(x) in self.LOOP_KW

        while i < len(toks):
            k = str(toks[i])
            if k in ("FOR", "AS"):
                var, how = toks[i + 1], str(toks[i + 2])
                if how == "IN":
                    iters.append(("in", var, toks[i + 3]))
                    i += 4
                elif how == "FROM":
                    start, end, by, below = toks[i + 3], None, None, False
                    i += 4
                    while i < len(toks) and isinstance(toks[i], Sym) and str(toks[i]) in ("TO", "BELOW", "BY", "UPTO"):
                        w = str(toks[i])
                        if w == "BY":
                            by = toks[i + 1]
                        else:
                            end, below = toks[i + 1], (w == "BELOW")
                        i += 2
                    iters.append(("from", var, start, end, by, below))
                elif how == "=":
                    then = None
                    expr = toks[i + 3]
                    i += 4
                    if i < len(toks) and isinstance(toks[i], Sym) and str(toks[i]) == "THEN":
                        then = toks[i + 1]
                        i += 2
                    iters.append(("=", var, expr, then))
                elif how == "BEING":
                    # for k being the hash-keys of table
                    which = str(toks[i + 4])
                    iters.append(("hash", var, which, toks[i + 6]))
                    i += 7
                else:
                    raise LispError("loop for " + how)
            elif k == "REPEAT":
                iters.append(("repeat", None, toks[i + 1]))
                i += 2
            elif k in ("WHILE", "UNTIL", "COLLECT", "APPEND", "SUM"):
                body_clauses.append((k, toks[i + 1]))
                i += 2
            elif k == "DO":
                j = i + 1
                forms = []
                while j < len(toks) and not kw(toks[j]):
                    forms.append(toks[j])
                    j += 1
                body_clauses.append(("DO", forms))
                i = j
            else:
                raise LispError("loop keyword " + k)

        new = Env(env)
        state = []
        for it in iters:
            kind = it[0]
            if kind == "in":
                state.append(["in", it[1], list(self.eval(it[2], new, fenv) or []), 0])
                new.vars[it[1]] = None
            elif kind == "from":
                start = self.eval(it[2], new, fenv)
                end = self.eval(it[3], new, fenv) if it[3] is not None else None
                by = self.eval(it[4], new, fenv) if it[4] is not None else 1
                state.append(["from", it[1], start, end, by, it[5]])
                new.vars[it[1]] = start
            elif kind == "=":
                state.append(["=", it[1], it[2], it[3], True])
                new.vars[it[1]] = None
            elif kind == "hash":
                table = self.eval(it[3], new, fenv)
                items = [kv[0] if "KEY" in it[2] else kv[1] for kv in table.d.values()]
                state.append(["in", it[1], items, 0])
                new.vars[it[1]] = None
            elif kind == "repeat":
                state.append(["repeat", None, self.eval(it[2], new, fenv), 0])
        collected, total, mode = [], 0, None
        while True:
            stop = False
            for st in state:
                kind = st[0]
                if kind == "in":
                    if st[3] >= len(st[2]):
                        stop = True
                        break
                    new.vars[st[1]] = st[2][st[3]]
                    st[3] += 1
                elif kind == "from":
                    cur = st[2]
                    if st[3] is not None and (cur >= st[3] if st[5] else cur > st[3]):
                        stop = True
                        break
                    new.vars[st[1]] = cur
                    st[2] = cur + st[4]
                elif kind == "=":
                    if st[4] or st[3] is None:
                        new.vars[st[1]] = self.eval(st[2], new, fenv)
                        st[4] = False
                    else:
                        new.vars[st[1]] = self.eval(st[3], new, fenv)
                elif kind == "repeat":
                    if st[3] >= st[2]:
                        stop = True
                        break
                    st[3] += 1
            if stop:
                break
            done = False
            for k, arg in body_clauses:
                if k == "WHILE":
                    if not truthy(self.eval(arg, new, fenv)):
                        done = True
                        break
                elif k == "UNTIL":
                    if truthy(self.eval(arg, new, fenv)):
                        done = True
                        break
                elif k == "COLLECT":
                    collected.append(self.eval(arg, new, fenv))
                    mode = "list"
                elif k == "APPEND":
                    collected.extend(self.eval(arg, new, fenv) or [])
                    mode = "list"
                elif k == "SUM":
                    total += self.eval(arg, new, fenv)
                    mode = "sum"
                else:
                    for form in arg:
                        self.eval(form, new, fenv)
            if done:
                break
            if not state and not any(k in ("WHILE", "UNTIL") for k, _ in body_clauses):
                raise LispError("loop without an iteration or termination clause")
        if mode == "sum":
            return total
        return lst(collected)


class Generic:
    def __init__(self, interp, name):
        self.interp, self.name = interp, name
        self.methods = []          # (specializers, fn)
        self.readers = {}          # class name -> slot
        this = self

        def call(*args):
            return this.dispatch(*args)
        self.fn = call

    def add_reader(self, cname, slot):
        self.readers[cname] = slot

    def accessor_slot(self, obj):
        return self.readers.get(obj.cls.name)

    def add_method(self, specs, fn):
        for i, (s, _) in enumerate(self.methods):
            if s == specs:
                self.methods[i] = (specs, fn)
                return
        self.methods.append((specs, fn))

    def has_method_for(self, cname):
        return any(s and s[0] == cname for s, _ in self.methods)

    def dispatch(self, *args):
        it = self.interp
        best, best_score = None, -1
        for specs, fn in self.methods:
            if len(args) < len(specs):
                continue
            score = 0
            ok = True
            for a, s in zip(args, specs):
                if s is None or str(s) == "T":
                    continue
                if it.class_of(a) == s:
                    score += 1
                else:
                    ok = False
                    break
            if ok and score > best_score:
                best, best_score = fn, score
        if best is not None:
            return best(*args)
        if args and isinstance(args[0], Instance):
            slot = self.readers.get(args[0].cls.name)
            if slot is not None and len(args) == 1:
                try:
                    return args[0].slots[slot]
                except KeyError:
                    raise LispError("unbound slot %s in %s" % (slot, args[0].cls.name))
        raise LispError("no applicable method for %s on %s" % (self.name, [str(it.class_of(a)) for a in args]))


# ---------------------------------------------------------------------------
# built-in functions
# ---------------------------------------------------------------------------
def kwargs(args, start):
    out = {}
    for k in range(start, len(args) - 1, 2):
        out[str(args[k])] = args[k + 1]
    return out


def install_builtins(it: Interp):
    g = it.gfun

    def defb(name, fn):
        g[S(name)] = fn

    def num_cmp(op):
        def f(*a):
            for x in a:
                if not isinstance(x, (int, Fraction)) or x is True or x is None:
                    raise LispError("not a number: " + lisp_repr(x))
            return True if all(op(a[i], a[i + 1]) for i in range(len(a) - 1)) else None
        return f

    def chk(*a):
        for x in a:
            if not isinstance(x, (int, Fraction)) or isinstance(x, bool):
                raise LispError("not a number: " + lisp_repr(x))

    def add(*a):
        chk(*a)
        return sum(a)

    def mul(*a):
        chk(*a)
        out = 1
        for x in a:
            out *= x
        return out

    def sub(*a):
        chk(*a)
        if len(a) == 1:
            return -a[0]
        out = a[0]
        for x in a[1:]:
            out -= x
        return out

    def div(*a):
        chk(*a)
        out = Fraction(a[0])
        for x in a[1:]:
            out /= x
        return int(out) if out.denominator == 1 else out

    defb("+", add)
    defb("*", mul)
    defb("-", sub)
    defb("/", div)
    import operator as op
    defb("<", num_cmp(op.lt))
    defb("<=", num_cmp(op.le))
    defb(">", num_cmp(op.gt))
    defb(">=", num_cmp(op.ge))
    defb("=", num_cmp(op.eq))
    defb("MAX", lambda *a: (chk(*a), max(a))[1])
    defb("MIN", lambda *a: (chk(*a), min(a))[1])
    defb("MOD", lambda a, b: (chk(a, b), a % b)[1])
    defb("FLOOR", lambda a, b=1: (chk(a, b), a // b)[1])
    defb("ABS", lambda a: abs(a))
    defb("1+", lambda a: a + 1)
    defb("1-", lambda a: a - 1)
    defb("NUMERATOR", lambda a: Fraction(a).numerator)
    defb("DENOMINATOR", lambda a: Fraction(a).denominator)
    defb("RANDOM", lambda n: it.rng.randrange(n))
    defb("NOT", lambda x: True if not truthy(x) else None)
    defb("NULL", lambda x: True if not truthy(x) else None)
    defb("EQ", lambda a, b: True if eq(a, b) else None)
    defb("EQL", lambda a, b: True if eq(a, b) else None)
    defb("EQUAL", lambda a, b: True if equal(a, b) else None)
    defb("STRING=", lambda a, b: True if str(a) == str(b) else None)
    defb("IDENTITY", lambda x: x)
    defb("LISTP", lambda x: True if (x is None or isinstance(x, list)) else None)
    defb("CONSP", lambda x: True if isinstance(x, list) else None)
    defb("NUMBERP", lambda x: True if (isinstance(x, (int, Fraction)) and not isinstance(x, bool)) else None)
    defb("SYMBOLP", lambda x: True if (x is None or x is True or isinstance(x, Sym)) else None)
    defb("STRINGP", lambda x: True if (isinstance(x, str) and not isinstance(x, (Sym, Char))) else None)
    defb("FUNCTIONP", lambda x: True if (callable(x) and not isinstance(x, str)) else None)

    def type_of(x):
        if isinstance(x, Instance):
            return x.cls.name
        if isinstance(x, Fraction):
            return S("RATIO")
        if isinstance(x, bool) or x is None or isinstance(x, Sym):
            return S("SYMBOL") if x is not None else S("NULL")
        if isinstance(x, int):
            return S("INTEGER")
        if isinstance(x, list):
            return S("CONS")
        if isinstance(x, str):
            return S("STRING")
        return S("FUNCTION")
    defb("TYPE-OF", type_of)

    def need_list(x):
        if x is not None and not isinstance(x, list):
            raise LispError("not a list: " + lisp_repr(x))
        return x or []

    def nth(i, l):
        if not isinstance(i, int) or isinstance(i, bool) or i < 0:
            raise LispError("nth: bad index " + lisp_repr(i))
        l = need_list(l)
        return l[i] if i < len(l) else None

    defb("CAR", lambda l: nth(0, l))
    defb("FIRST", lambda l: nth(0, l))
    defb("SECOND", lambda l: nth(1, l))
    defb("THIRD", lambda l: nth(2, l))
    defb("FOURTH", lambda l: nth(3, l))
    defb("NTH", nth)
    def cdr(l):
        l = need_list(l)
        if isinstance(l, Dotted):
            return l.tail if len(l) == 1 else Dotted(l[1:], l.tail)
        return lst(l[1:])
    defb("CDR", cdr)
    defb("REST", cdr)

    def assoc(item, alist, *rest):
        t = testfn(kwargs(rest, 0))
        for e in need_list(alist):
            if e is not None and t(item, e[0]):
                return e
        return None
    defb("ASSOC", assoc)

    def sublis(alist, tree):
        def walk(x):
            if isinstance(x, list):
                out = [walk(e) for e in x]
                return Dotted(out, walk(x.tail)) if isinstance(x, Dotted) else out
            if isinstance(x, Sym) or x is True:
                for e in need_list(alist):
                    if eq(e[0], x):
                        return cdr(e)
            return x
        return walk(tree)
    defb("SUBLIS", sublis)
    defb("STRING-DOWNCASE", lambda x: LStr(str(x).lower()))
    defb("STRING-UPCASE", lambda x: LStr(str(x).upper()))
    defb("NTHCDR", lambda n, l: lst(need_list(l)[n:]))
    defb("LAST", lambda l: lst(need_list(l)[-1:]))
    defb("LIST", lambda *a: lst(list(a)))

    def cons(a, b):
        if b is not None and not isinstance(b, list):
            raise LispError("dotted pair not supported: " + lisp_repr(b))
        return [a] + (b or [])
    defb("CONS", cons)

    def append(*ls):
        out = []
        for l in ls:
            out.extend(need_list(l))
        return lst(out)
    defb("APPEND", append)
    defb("COPY-LIST", lambda l: lst(list(need_list(l))))
    defb("REVERSE", lambda l: lst(list(reversed(need_list(l)))) if not isinstance(l, str) else LStr(l[::-1]))

    def length(x):
        if x is None:
            return 0
        if isinstance(x, (list, str)):
            return len(x)
        raise LispError("length of " + lisp_repr(x))
    defb("LENGTH", length)

    def subseq(x, a, b=None):
        if isinstance(x, str):
            if a > len(x) or (b is not None and b > len(x)):
                raise LispError("subseq out of range")
            return LStr(x[a:b])
        x = need_list(x)
        if a > len(x) or (b is not None and (b > len(x) or b < a)):
            raise LispError("subseq out of range")
        return lst(x[a:b])
    defb("SUBSEQ", subseq)

    def funcall(f, *a):
        if isinstance(f, Sym):
            f = it.fn(f)
        return f(*a)

    def apply(f, *a):
        if isinstance(f, Sym):
            f = it.fn(f)
        return f(*(list(a[:-1]) + need_list(a[-1])))
    defb("FUNCALL", funcall)
    defb("APPLY", apply)

    def mapcar(f, *ls):
        if isinstance(f, Sym):
            f = it.fn(f)
        ls = [need_list(l) for l in ls]
        return lst([f(*xs) for xs in zip(*ls)])
    defb("MAPCAR", mapcar)

    def testfn(kw):
        t = kw.get(":TEST")
        key = kw.get(":KEY")
        if t is None:
            t = lambda a, b: eq(a, b)            # eql
        if key is None:
            return lambda item, e: truthy(t(item, e))
        return lambda item, e: truthy(t(item, key(e)))

    def position(item, seq, *rest):
        t = testfn(kwargs(rest, 0))
        if isinstance(seq, str):
            seq = [Char(c) for c in seq]
        for i, e in enumerate(need_list(seq)):
            if t(item, e):
                return i
        return None

    def find(item, seq, *rest):
        t = testfn(kwargs(rest, 0))
        for e in need_list(seq):
            if t(item, e):
                return e
        return None

    def position_if(pred, seq, *rest):
        for i, e in enumerate(need_list(seq)):
            if truthy(pred(e)):
                return i
        return None

    def find_if(pred, seq, *rest):
        for e in need_list(seq):
            if truthy(pred(e)):
                return e
        return None

    def remove(item, seq, *rest):
        t = testfn(kwargs(rest, 0))
        return lst([e for e in need_list(seq) if not t(item, e)])

    def remove_if(pred, seq, *rest):
        return lst([e for e in need_list(seq) if not truthy(pred(e))])

    def count(item, seq, *rest):
        t = testfn(kwargs(rest, 0))
        return sum(1 for e in need_list(seq) if t(item, e))
    defb("POSITION", position)
    defb("FIND", find)
    defb("POSITION-IF", position_if)
    defb("FIND-IF", find_if)
    defb("REMOVE", remove)
    defb("REMOVE-IF", remove_if)
    defb("COUNT", count)
    defb("MEMBER", lambda item, seq, *r: next((lst(need_list(seq)[i:]) for i, e in enumerate(need_list(seq))
                                                  if testfn(kwargs(r, 0))(item, e)), None))

    def sort(seq, pred, *rest):
        """SBCL sorts lists with a stable merge sort; only (pred later earlier) decides a swap."""
        key = kwargs(rest, 0).get(":KEY")
        items = list(need_list(seq))

        def less(a, b):
            return truthy(pred(key(a), key(b))) if key else truthy(pred(a, b))

        def msort(xs):
            if len(xs) <= 1:
                return xs
            h = len(xs) // 2
            a, b = msort(xs[:h]), msort(xs[h:])
            out, i, j = [], 0, 0
            while i < len(a) and j < len(b):
                if less(b[j], a[i]):
                    out.append(b[j])
                    j += 1
                else:
                    out.append(a[i])
                    i += 1
            return out + a[i:] + b[j:]
        return lst(msort(items))
    defb("SORT", sort)
    defb("STABLE-SORT", sort)

    def getf(plist, key, default=None):
        pl = need_list(plist)
        for k in range(0, len(pl) - 1, 2):
            if eq(pl[k], key):
                return pl[k + 1]
        return default
    defb("GETF", getf)
    defb("MAKE-HASH-TABLE", lambda *a: HashTable())

    def gethash(key, table, default=None):
        v = table.d.get(hash_key(key))
        return v[1] if v is not None else default
    defb("GETHASH", gethash)
    defb("MAKE-INSTANCE", lambda c, *a: it.make_instance(c, *a))

    def slot_value(obj, name):
        try:
            return obj.slots[name]
        except KeyError:
            raise LispError("unbound slot " + str(name))
    defb("SLOT-VALUE", slot_value)

    def error(*a):
        raise LispError("error: " + " ".join(lisp_repr(x) for x in a))
    defb("ERROR", error)

    def fmt(dest, ctrl, *a):
        s = lisp_format(ctrl, a)
        if dest is None:
            return s
        if isinstance(dest, StringStream):
            dest.parts.append(str(s))
            return None
        it.out.append(str(s))
        return None
    defb("FORMAT", fmt)
    defb("EVAL", lambda form: it.eval(keep(form), None, None))

    def keep(form):
        it._keep.append(form)       # expansions are cached by id(form): runtime-built forms must stay alive
        return form

    defb("READ-FROM-STRING", lambda s: read_from_string(str(s)))
    defb("CHAR", lambda s, i: Char(s[i]))

    def search(a, b):
        i = str(b).find(str(a))
        return i if i >= 0 else None
    defb("SEARCH", search)
    defb("GET-UNIVERSAL-TIME", lambda: 0)
    defb("SLEEP", lambda n: None)
    defb("PRINT", lambda x: x)


def load_file(it: Interp, path, first_line=0, last_line=10 ** 9, on_test=None, skip_heads=("LOAD", "QL:QUICKLOAD", "IN-PACKAGE"),
              only_definitions=True):
    """Evaluate the top-level forms of a reference source file.  Definitions are always evaluated; `(test F R P)`
    forms are handed to on_test(line, form) (or skipped); other top-level forms run only if only_definitions is false."""
    text = open(path, encoding="utf-8").read()
    defs = {"DEFUN", "DEFMACRO", "DEFCLASS", "DEFMETHOD", "DEFGENERIC", "DEFVAR", "DEFPARAMETER", "DEFINE-CONDITION"}
    for line, form in read_all_with_lines(text):
        if line < first_line or line > last_line or not isinstance(form, list):
            continue
        head = str(form[0])
        if head in skip_heads:
            continue
        if head == "TEST":
            if on_test is not None:
                on_test(line, form)
            continue
        if head in defs or not only_definitions:
            it.eval(form, None, None)
