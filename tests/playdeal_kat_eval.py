"""Evaluator for the inline test forms of the reference's trick estimator (bridge.cl:1497-2984).

The known-answer tests of `play-deal` and its parts are Lisp *code* (`(as-list (beat-or-low (make-instance 'trick
...) ...))`), not data, so tests/golden/make_golden.py stores them as text and this module evaluates them: a form
walker over the handful of special forms they use (quote, backquote, function, let, let*, lambda) whose function
table is supplied by the caller -- the oracle's Python restatement (tests/test_playdeal_oracle.py) or the host
mirror of the CUDA path (tests/test_gpu_playdeal.py).  Expected values are the reference's own, read from the
fixture."""
from bridge_mc_b200 import sexp

Sym = sexp.Sym


def norm(x):
    """Lisp value -> comparable Python value: nil and () are the same thing, symbols C D H S stay symbols."""
    if x is None or x is False:
        return []
    if isinstance(x, (list, tuple)):
        return [norm(e) for e in x]
    return x


def equal(a, b):
    return norm(a) == norm(b)


class Env(dict):
    def __init__(self, parent=None):
        super().__init__()
        self.parent = parent

    def lookup(self, k):
        e = self
        while e is not None:
            if k in e:
                return e[k]
            e = e.parent
        raise KeyError(k)


def is_kw(x):
    return isinstance(x, Sym) and x.startswith(":")


def split_args(args):
    """positional values and keyword values of an evaluated argument list"""
    pos, kw = [], {}
    i = 0
    while i < len(args):
        if is_kw(args[i]) and i + 1 < len(args):
            kw[str(args[i])[1:].lower().replace("-", "_")] = args[i + 1]
            i += 2
        else:
            pos.append(args[i])
            i += 1
    return pos, kw


class Evaluator:
    def __init__(self, functions):
        """functions: {LISP-NAME: python callable(*positional, **keywords)}"""
        self.fn = dict(functions)

    def function(self, name):
        if isinstance(name, list) and name and name[0] == "LAMBDA":
            raise NotImplementedError
        return self.fn[str(name)]

    def quasi(self, x, env):
        if isinstance(x, list):
            if len(x) == 2 and x[0] == "UNQUOTE":
                return self.eval(x[1], env)
            out = []
            for e in x:
                if isinstance(e, list) and len(e) == 2 and e[0] == "UNQUOTE-SPLICE":
                    out.extend(self.eval(e[1], env) or [])
                else:
                    out.append(self.quasi(e, env))
            return out
        return x

    def eval(self, f, env=None):
        env = env if env is not None else Env()
        if isinstance(f, Sym):
            if is_kw(f):
                return f
            return env.lookup(str(f))
        if not isinstance(f, list):
            return f
        if not f:
            return None
        h = f[0]
        if h == "QUOTE":
            return f[1]
        if h == "BACKQUOTE":
            return self.quasi(f[1], env)
        if h == "FUNCTION":
            if isinstance(f[1], list):
                return self.eval(f[1], env)
            return self.function(f[1])
        if h in ("LET", "LET*"):
            new = Env(env)
            for name, val in f[1]:
                new[str(name)] = self.eval(val, new if h == "LET*" else env)
            out = None
            for body in f[2:]:
                out = self.eval(body, new)
            return out
        if h == "LAMBDA":
            params = [str(p) for p in f[1]]
            body = f[2:]

            def closure(*args):
                new = Env(env)
                for p, a in zip(params, args):
                    new[p] = a
                out = None
                for b in body:
                    out = self.eval(b, new)
                return out
            return closure
        if h == "PEEK":
            return self.eval(f[1], env)
        args = [self.eval(a, env) for a in f[1:]]
        if h == "APPLY":
            fn, rest = args[0], list(args[1:-1]) + list(args[-1] or [])
            pos, kw = split_args(rest)
            return fn(*pos, **kw)
        if h == "CURRY":
            fn, base = args[0], args[1:]

            def curried(*more):
                pos, kw = split_args(list(base) + list(more))
                return fn(*pos, **kw)
            return curried
        if h == "LIST":
            return list(args)
        if h == "CONS":
            return [args[0]] + list(args[1] or [])
        if h == "APPEND":
            out = []
            for a in args:
                out.extend(a or [])
            return out
        if h == "MAPCAR":
            return [args[0](x) for x in (args[1] or [])]
        pos, kw = split_args(args)
        return self.fn[str(h)](*pos, **kw)
