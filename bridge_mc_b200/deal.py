"""Host mirror of the reference's deal generator and Monte-Carlo driver:
class `deal` / `build` (bridge.cl:1283-1319), class `mc-case` with sim / select /
choose / add-prop / save / load-mc-case (bridge.cl:3133-3221) and the REST
entry `mc-deal` with normdef / invert-id / close-list (rest.cl:214-277).

`build` keeps its signature and result order (const hands, `:~` seats in def
order, then the random hands) but every deal comes from the GPU sampler
(bmc_sample): one library call produces a batch that successive `build`s
consume, and `sim` asks for all its deals in one call.
"""
from __future__ import annotations

import itertools
import os

import numpy as np

from . import hist as _hist
from . import sexp
from .cards import Hand, str2hand, unpack_deal, seat_masks
from .engine import Constraints, DEFAULT_SEED, Engine, seat_spec
from .infer import infer_distgen_params, plist_to_def, supply
from .lists import roll  # noqa: F401  (bridge.cl:420-425; re-exported)


class NoDealFound(RuntimeError):
    """The raw-deal budget was used up without a single accepted deal."""


def normdef(d):
    """rest.cl:250-256: number -> (n n); (lo hi n) -> (lo hi (n n))."""
    if isinstance(d, dict):
        return {k: _normval(v) for k, v in d.items()}
    return [_normval(v) if i % 2 else v for i, v in enumerate(d)]


def _normval(v):
    if isinstance(v, int) and not isinstance(v, bool):
        return [v, v]
    if isinstance(v, (list, tuple)) and len(v) > 2 and isinstance(v[2], int):
        return [v[0], v[1], [v[2], v[2]]]
    return v


def invert_id(ids):
    """rest.cl:258-259: inverse permutation."""
    return [i for i, _ in sorted(enumerate(ids), key=lambda p: p[1])]


def close_list(lst, base):
    """rest.cl:214-217"""
    return list(lst) + list(base[len(lst):]) if len(lst) < len(base) else list(lst)


def reorder(lst, indexes):
    """bridge.cl:202-204"""
    return [lst[i] if i < len(lst) else None for i in indexes]


class Deal:
    """(make-instance 'deal := const-hands :~ defs), bridge.cl:1283-1286.

    const_hands: Hands (or 4-list rank lists / "♣ .. ♦ .." strings); defs: range
    defs (dict, plist as read by sexp, or plist text); rules: optional per-`:~`-seat
    rule forms in the bidding.cl predicate language (an extension: the reference
    has no entry point that feeds bidding predicates to the generator)."""

    BATCH = 2048

    def __init__(self, const_hands=None, defs=None, rules=None, engine: Engine | None = None, seed: int | None = None,
                 max_raw: int = 1 << 34, semantics: str | None = None):
        self.const_hands = [self._hand(h) for h in (const_hands or [])]
        self.defs = [plist_to_def(sexp.read(d) if isinstance(d, str) else d) for d in (defs or [])]
        self.rules = list(rules or [])
        if len(self.const_hands) + len(self.defs) > 4:
            raise ValueError("more than four seats defined")
        # "reference": `build`'s own sequential seat-by-seat sampling incl. its quirks (bmc_constraints_set_reference_order),
        # the default for plain range defs; "uniform": uniform over the deals satisfying every seat (joint rejection),
        # forced when rule forms are given (the reference generator cannot take them).
        self.semantics = semantics or ("uniform" if any(self.rules) else "reference")
        if self.semantics not in ("reference", "uniform"):
            raise ValueError("semantics must be 'reference' or 'uniform'")
        if self.semantics == "reference" and any(self.rules):
            raise ValueError("rule forms need semantics='uniform'")
        self._engine = engine
        self.seed = DEFAULT_SEED ^ int.from_bytes(os.urandom(8), "little") if seed is None else seed
        self.max_raw = max_raw
        self._next_index = 0
        self._buffer = []
        self._carry = None
        self._cons = None

    @staticmethod
    def _hand(h):
        if isinstance(h, Hand):
            return h
        if isinstance(h, str):
            return str2hand(h)
        return Hand(h)

    def __repr__(self):
        return f"DEAL<{self.const_hands}, {self.defs}>"

    @property
    def engine(self):
        return self._engine or Engine.default()

    def constraints(self) -> Constraints:
        if self._cons is None:
            # the reference narrows the first def against the others when the root
            # generator is made (bridge.cl:1297-1301) and signals bad-range there
            used = set(itertools.chain.from_iterable(h.cards() for h in self.const_hands))
            if self.defs:
                infer_distgen_params(supply(c for c in range(52) if c not in used), self.defs[0], self.defs[1:])
            specs = [seat_spec(fixed=h) for h in self.const_hands]
            specs += [seat_spec(**{k: v for k, v in d.items() if k in ("hcp", "c", "d", "h", "s", "losers")}) for d in self.defs]
            specs += [None] * (4 - len(specs))
            rules = [None] * len(self.const_hands) + list(self.rules)
            rules += [None] * (4 - len(rules))
            self._cons = Constraints(specs, rules if any(rules) else None)
            if self.semantics == "reference":
                # no def at all still samples one seat through the root generator (bridge.cl:1297-1313)
                n_defs = max(1, len(self.defs))
                if len(self.const_hands) + n_defs <= 4:
                    self._cons.set_reference_order(n_defs)
        return self._cons

    def _shuffled(self, rec: np.ndarray, first_index: int) -> np.ndarray:
        """The accepted SET of a library call is fixed by (seed, index range); the order in which the kernel stored it is
        not.  Sorting makes the batch reproducible, and a permutation that depends on (seed, first_index) only -- never on
        the cards -- makes every prefix of it an unbiased sample, so callers may take fewer deals than a call produced."""
        rec = rec[np.lexsort((rec[:, 0], rec[:, 1]))]
        rng = np.random.default_rng([self.seed & 0xFFFFFFFFFFFFFFFF, first_index & 0xFFFFFFFFFFFFFFFF])
        return rec[rng.permutation(rec.shape[0])]

    def build_records(self, n: int) -> np.ndarray:
        """n accepted deals as uint64[n,2] records (seat k = k-th of const, `:~`, random).  Deals a library call produced
        beyond `n` are kept for the next call (the kernels run whole waves)."""
        cons = self.constraints()
        out = []
        have = 0
        if self._carry is not None and self._carry.shape[0]:
            take = self._carry[:n]
            self._carry = self._carry[take.shape[0]:]
            out.append(take)
            have = take.shape[0]
        every_index_deals = self.semantics == "reference" and cons.mode in (0, 3)
        wave = 1 << 14
        spent = 0
        while have < n:
            first = self._next_index
            if every_index_deals:
                # reference order: one deal per index (unless the reference would signal an error for it) -- ask for
                # exactly the missing ones, nothing is overdrawn
                need = n - have
                rec, tal, st = self.engine.sample(cons, n_raw=need, seed=self.seed, first_index=first, tally_mask=0, cap=need)
            else:
                budget = max(wave, min(self.max_raw - spent, wave * 256))          # bounded work per library call
                rec, tal, st = self.engine.sample(cons, n_accept=n - have, n_raw=budget, seed=self.seed, first_index=first,
                                                  wave=wave, tally_mask=0, cap=(n - have) + wave)
            used = int(st.raw)
            self._next_index += used
            spent += used
            if rec.shape[0]:
                rec = self._shuffled(rec, first)
                out.append(rec)
                have += rec.shape[0]
            elif cons.mode == 3 and st.voided == st.raw:
                from .infer import BadRangeError
                raise BadRangeError("Can't infer correct limit!", None, None)      # every build signals bad-range
            elif spent >= self.max_raw:
                raise NoDealFound(f"no deal satisfies {self!r} within {spent} raw deals")
            if not every_index_deals and st.accepted * 64 < used and wave < (1 << 30):
                wave <<= 2                                   # rare constraint: take bigger bites
        rec = np.concatenate(out) if out else np.empty((0, 2), np.uint64)
        if rec.shape[0] > n:
            self._carry = rec[n:]
        return rec[:n]

    def build(self):
        """bridge.cl:1292-1319: one deal as a list of Hands -- const hands, `:~` seats, random hands.
        With three const hands and one def only the sampled hand is returned (bridge.cl:1316-1319)."""
        if not self._buffer:
            recs = self.build_records(self.BATCH if not self.const_hands else 256)
            self._buffer = [recs[i] for i in range(recs.shape[0])][::-1]
        hands = unpack_deal(self._buffer.pop(), ids=None)
        if len(self.const_hands) == 3 and len(self.defs) == 1:
            return hands[3:4]
        return hands


class McCase:
    """bridge.cl:3133-3221 class mc-case.  `dealgen` is a Deal (optionally with a
    seat permutation `reorder`) or any callable returning four Hands; `props` are
    measures called as measure(trump, n, e, s, w) (README.md:129-133)."""

    def __init__(self, dealgen, props, trump=None, data=None, reorder=None):
        self.dealgen, self.props, self.trump = dealgen, list(props), trump
        self.data = list(data or [])
        self.reorder = reorder
        self._threads = []

    def __repr__(self):
        return f"MC-CASE{{{self.dealgen}->{[getattr(p, '__name__', p) for p in self.props]} @ {self.trump}}}"

    def _deals(self, runs):
        if isinstance(self.dealgen, Deal):
            recs = self.dealgen.build_records(runs)
            for i in range(recs.shape[0]):
                hands = unpack_deal(recs[i], ids=None)
                yield reorder(hands, self.reorder) if self.reorder else hands
        else:
            for _ in range(runs):
                yield self.dealgen()

    def sim(self, runs: int):
        """bridge.cl:3144-3162: `runs` rows (hands m1 m2 ...) appended to the case's data.
        The reference starts two threads and returns at once; here the deals come
        from one batched GPU call and the rows are complete on return."""
        rows = [[hands] + [p(self.trump, *hands) for p in self.props] for hands in self._deals(runs)]
        self.data.extend(rows)
        return rows

    def threads(self):
        """bridge.cl:3164-3167: live worker threads (always none: sim is synchronous)."""
        return []

    def add_prop(self, *new_props):
        """bridge.cl:3169-3175"""
        for row in self.data:
            row.extend(p(self.trump, *row[0]) for p in new_props)
        self.props.extend(new_props)

    def _names(self):
        return ["hands"] + [getattr(p, "__name__", str(p)) for p in self.props]

    def choose(self, pred):
        """bridge.cl:3177-3184: rows for which pred(row_dict) holds (row_dict maps 'hands' and
        each measure name to its column; the reference `eval`s a form over the same names)."""
        names = self._names()
        return [row for row in self.data if pred(dict(zip(names, row)))]

    def select(self, fields=None, filter=None):
        """bridge.cl:3186-3195"""
        source = self.choose(filter) if filter else self.data
        if fields:
            names = self._names()
            cols = [names.index(f) for f in fields]
            return [[row[c] for c in cols] for row in source]
        return [row[1:] for row in source]

    def save(self, filename):
        """bridge.cl:3201-3209.  Hands are written as records (lo hi), which -- unlike the
        reference's HAND<..> print form -- read back exactly."""
        from .cards import pack_deal
        with open(filename, "w", encoding="utf-8") as f:
            for row in self.data:
                rec = pack_deal(row[0])
                f.write(sexp.dumps([[int(rec[0]), int(rec[1])]] + [_to_sexp(v) for v in row[1:]]) + "\n")

    @classmethod
    def load(cls, filename, fallback=None, sim_count=None):
        """bridge.cl:3211-3221 load-mc-case"""
        if os.path.exists(filename):
            rows = []
            for form in sexp.read_all(open(filename, encoding="utf-8").read()):
                rec = np.array(form[0], dtype=np.uint64)
                rows.append([unpack_deal(rec, ids=None)] + list(form[1:]))
            kw = dict(fallback or {})
            kw.setdefault("dealgen", None)
            kw.setdefault("props", [])
            return cls(data=rows, **kw)
        if fallback:
            obj = cls(**fallback)
            if sim_count:
                obj.sim(sim_count)
            return obj
        return None


def _to_sexp(v):
    if isinstance(v, (list, tuple)):
        return [_to_sexp(x) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    return v


def mc_deal(defs, measures, runs: int = 500, engine: Engine | None = None, seed: int | None = None):
    """rest.cl:264-277 mc-deal: `defs` in the request's seat order (N E S W), each a Hand
    (fixed) or a range def; builds the `deal`, the seat permutation back to request
    order, simulates `runs` deals (the reference hard-codes 500) and returns
    (mc_case, histogram of the measure tuples)."""
    consts = [(i, d) for i, d in enumerate(defs) if isinstance(d, Hand)]
    ranged = [(i, d) for i, d in enumerate(defs) if not isinstance(d, Hand)]
    deal = Deal([d for _, d in consts], [normdef(plist_to_def(d)) for _, d in ranged], engine=engine, seed=seed)
    perm = close_list(invert_id([i for i, _ in consts + ranged]), [0, 1, 2, 3])
    mc = McCase(deal, measures, reorder=perm)
    mc.sim(runs)
    rows = [list(itertools.chain.from_iterable(v if isinstance(v, list) else [v] for v in r)) for r in mc.select()]
    return mc, _hist.histogram(rows)


def hand_or_def_parser(defs):
    """rest.cl:279-285 hand-or-def-parser: the request's `defs`, a list of (type, text) pairs; type "hand" gives a
    fixed hand (str2hand), anything else is the printed range def, e.g. "(:hcp (15 17) :s (5 nil))", read as Lisp
    reads it.  The result is what mc_deal takes."""
    out = []
    for kind, val in defs:
        if str(kind).lstrip(":").strip("|").lower() == "hand":
            out.append(str2hand(val))
        else:
            out.append(plist_to_def(sexp.read(val)))
    return out


def measure_parser(names, best_tricks=None, engine: Engine | None = None):
    """rest.cl:287-303 measure-parser: contract names ("3NTS", "4HSx") -> measures `msr-<name>` returning
    [tricks, base-score of the contract with that many tricks].  The tricks come from `best-tricks`, i.e. the
    `play-deal` estimator (bridge.cl:2961), which runs on the GPU (bridge_mc_b200/playdeal.py, bmc_play_hands);
    `best_tricks(suit_sym, declarer_hand, left, dummy, right) -> int` replaces it if given.  [""] selects the
    reference's default measure `best-outcomes` (bridge.cl:3071)."""
    from . import compscore, playdeal
    if best_tricks is None:
        def best_tricks(suit, *hands):
            return playdeal.best_tricks(suit, *hands, engine=engine)
    if list(names) == [""]:
        def best_outcomes(_trump, *hands):
            return playdeal.best_outcomes(None, *hands, engine=engine)
        return [best_outcomes]
    measures = []
    for name in names:
        contract = compscore.Contract(name)
        shift = "NESW".index(contract.declarer)

        def msr(_trump, *hands, _c=contract, _shift=shift, _name=name):
            tricks = int(best_tricks(_c.suit_sym(), *roll(-_shift, list(hands))))
            scored = compscore.Contract(_name).with_score(tricks)
            return [tricks, compscore.base_score(scored, engine=engine)]

        msr.__name__ = f"msr-{name}"
        measures.append(msr)
    return measures

