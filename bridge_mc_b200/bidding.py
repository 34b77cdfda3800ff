"""Host mirror of bidding.cl: rule tables and the entry points that evaluate them.

Rule tables are alists `((bid . form) ...)` in the reference's own predicate
language (bidding.cl:60-74); here they are built as S-expression text and handed
to libbridgemc (bmc_scheme_create / bmc_constraints_create), which compiles them
to device bytecode.  `assess_hand`, `test_bid` and `choose_bid` keep the
reference signatures (bidding.cl:33,111,146) and run on the GPU via bmc_filter.

Scheme *generators* (nt_responses, basic_responses) are host-side, as in the
reference: they only emit forms.
"""
from __future__ import annotations

import numpy as np

from . import sexp
from .cards import Hand, pack_deal, LANE_MASK
from .engine import Engine, Scheme, Constraints

SUITS = ("C", "D", "H", "S")


# ---- small form builders -------------------------------------------------------
def _and(*xs):
    return "(and " + " ".join(xs) + ")"


def _or(*xs):
    return "(or " + " ".join(xs) + ")"


def _not(x):
    return f"(not {x})"


def _cmp(op, a, b):
    return f"({op} {a} {b})"


def _any_longer_than(n, lo, hi):
    """some suit of lens[lo:hi] longer than n -- the reference's find-if/subseq idiom"""
    return f"(find-if (curry #'< {n}) (subseq lens {lo} {hi}))"


def _between(var, lo, hi):
    """bidding.cl:163-164 `range`"""
    return _and(_cmp(">=", var, lo), _cmp("<=", var, hi))


class RuleTable:
    """An ordered alist of (bid, form-text) rows; `bid` is a tuple like (1, 'NT')."""

    def __init__(self, rows):
        self.rows = [((tuple(b) if b is not None else None), f) for b, f in rows]     # bid None: a NIL row, never matches
        self._scheme = None

    def bids(self):
        return [b for b, _ in self.rows]

    def index(self, bid):
        bid = _norm_bid(bid)
        for i, (b, _) in enumerate(self.rows):
            if b == bid:
                return i
        raise KeyError(bid)

    def form(self, bid) -> str:
        return self.rows[self.index(bid)][1]

    def sexpr(self) -> str:
        return "(" + " ".join("nil" if b is None else f"(({b[0]} {b[1]}) . {f})" for b, f in self.rows) + ")"

    def scheme(self) -> Scheme:
        if self._scheme is None:
            self._scheme = Scheme(self.sexpr())
        return self._scheme

    # -- forms for seat constraints ------------------------------------------
    def chooses(self, bid) -> str:
        """form that is true iff choose-bid over this table returns `bid` (first match wins)"""
        i = self.index(bid)
        earlier = [f for _, f in self.rows[:i]]
        if not earlier:
            return self.rows[i][1]
        return _and(self.rows[i][1], _not(_or(*earlier)))

    def passes(self) -> str:
        """form that is true iff choose-bid over this table returns nil"""
        return _not(_or(*[f for _, f in self.rows]))


def _norm_bid(bid):
    if isinstance(bid, str):
        bid = sexp.read(bid)
    level, suit = bid
    return (int(level), str(suit).upper())


# ---- the tables (semantics: bidding.cl:76-106, 166-187, 204-210, 221-253) --------
def _opening_rows():
    rows = []
    sound = _or(_cmp("<", "loosers", 6), _cmp(">", "hcp", 11))
    # strong two clubs: 23+ or an unbalanced hand short of losers with a five-card suit
    rows.append(((2, "C"), _or(_cmp(">", "hcp", 22),
                               _and(_not("balanced"),
                                    "(cond (" + _any_longer_than(4, 2, 4) + " " + _cmp("<=", "loosers", 4) + ") "
                                    "(" + _any_longer_than(4, 0, 2) + " " + _cmp("<=", "loosers", 3) + "))"))))
    rows.append(((2, "NT"), _and("balanced", _not("(find-if (curry #'> 3) power)"), _cmp(">", "hcp", 18))))
    rows.append(((1, "NT"), _and("balanced", _cmp(">", "hcp", 14), _cmp("<", "hcp", 18))))
    for suit, upto in (("S", 3), ("H", 2)):      # a longer lower suit with 17+ bids that suit first (reverse)
        rows.append(((1, suit), _and(sound, _cmp(">", "loosers", 4), _cmp("<", "hcp", 23), _cmp(">=", suit, 5),
                                     _not(_and(_cmp(">", "hcp", 16), _any_longer_than(4, 0, upto))))))
    rows.append(((1, "D"), _and(sound, _cmp(">", "loosers", 3), _cmp("<", "hcp", 23), _cmp(">=", "D", 4),
                                _cmp(">=", "D", "C"), _not(_and(_cmp(">", "hcp", 16), _cmp(">=", "C", 5))))))
    rows.append(((1, "C"), _and(sound, _cmp(">", "loosers", 3), _cmp("<", "hcp", 23), _cmp(">=", "C", 3))))
    for suit in ("D", "H", "S"):                 # weak twos
        rows.append(((2, suit), _and(_cmp(">", "hcp", 5), _cmp(">=", suit + "power", 5),
                                     _or(_cmp(">", "loosers", 5), _cmp("<", "hcp", 11)), _cmp("=", suit, 6))))
    for suit in SUITS:                           # three-level preempts
        rows.append(((3, suit), _and(_cmp(">", "hcp", 5), _cmp(">=", suit + "power", 5), _cmp("<", "hcp", 11),
                                     _cmp(">=", suit, 7))))
    return rows


openings = RuleTable(_opening_rows())


def nt_responses(level: int) -> RuleTable:
    """bidding.cl:166-187: responses to a 1NT (level 1) or 2NT (level 2) opening."""
    invite = (8, 9) if level == 1 else (4, 5)
    game = (10, 15) if level == 1 else (6, 10)
    slam_invite = (16, 17) if level == 1 else (12, 13)
    slam = 18 if level == 1 else 14
    base = level + 1
    rows = [
        ((base, "C"), _and(_cmp(">=", "hcp", invite[0]), _or(_cmp("=", "S", 4), _cmp("=", "H", 4)))),   # Stayman
        ((base, "D"), _cmp(">=", "H", 5)),                                                              # transfers
        ((base, "H"), _cmp(">=", "S", 5)),
        ((base, "S"), _and(_or(_cmp("<", "hcp", invite[0]), _cmp(">", "hcp", game[1] - 2)),
                           _or(_and(_cmp(">", "hcp", invite[0]), _cmp(">=", "C", 5)), _cmp(">=", "C", 6)))),
    ]
    if level == 1:
        rows.append(((3, "C"), _and(_or(_cmp("<", "hcp", 8), _cmp(">", "hcp", 13)),
                                    _or(_and(_cmp(">=", "hcp", 7), _cmp(">=", "D", 5)), _cmp(">=", "D", 6)))))
        rows.append(((2, "NT"), _and(_cmp(">=", "hcp", 8), _cmp("<=", "hcp", 9), _cmp("<", "H", 5), _cmp("<", "S", 5),
                                     _cmp("<", "C", 6), _cmp("<", "D", 6))))
    rows += [((3, "NT"), _between("hcp", *game)),
             ((4, "NT"), _between("hcp", *slam_invite)),
             ((6, "NT"), _cmp(">=", "hcp", slam))]
    return RuleTable(rows)


two_c_responses = RuleTable([
    ((2, "D"), _cmp(">=", "hcp", 6)),
    ((2, "H"), _and(_cmp("<=", "hcp", 5), _cmp(">=", "H", 4))),
    ((2, "S"), _and(_cmp("<=", "hcp", 5), _cmp(">=", "S", 4))),
    ((2, "NT"), _and(_cmp("<=", "hcp", 5), "balanced")),
    ((3, "C"), _and(_cmp("<=", "hcp", 5), _cmp(">=", "C", 5), _not("balanced"))),
    ((3, "D"), _and(_cmp("<=", "hcp", 5), _cmp(">=", "D", 5), _not("balanced"))),
])


def fit(suit: str) -> str:
    """bidding.cl:221-224"""
    suit = suit.upper()
    if suit == "C":
        return _cmp(">=", "C", 5)
    if suit == "D":
        return _cmp(">=", "D", 4)
    return _cmp(">=", suit, 3)


def basic_responses(bid) -> RuleTable:
    """bidding.cl:226-253: responses to a one-of-a-suit opening.

    Kept quirk: in the `(3 suit)` raise the list of "no lower five-card suit"
    tests is inserted as ONE element instead of being spliced (bidding.cl:249),
    so that row is NIL for a club opening and an illegal call -- an error if
    evaluation reaches it -- for the other suits."""
    level, suit = _norm_bid(bid)
    if level > 1 or suit not in SUITS:
        raise ValueError("Wrong bid (only 1-suit bids accepted)")
    o = SUITS.index(suit)
    the_fit = fit(suit)
    no_biddable_4 = [_cmp("<", SUITS[i], 4) for i in range(o + 1, 4)]
    no_lower_5 = [_cmp("<", SUITS[i], 5) for i in range(0, o)]
    rows = [((1, SUITS[i]), _and(_cmp(">=", "hcp", 6), _cmp(">=", SUITS[i], 4))) for i in range(o + 1, 4)]
    rows.append(((1, "NT"), _and(_cmp(">=", "hcp", 6), _cmp("<=", "hcp", 10), *no_biddable_4, _not(the_fit))))
    if suit == "C":
        rows.append(((2, "C"), _and(_cmp(">=", "hcp", 6), _cmp("<=", "hcp", 9), _cmp(">=", "C", 5), *no_biddable_4)))
    else:
        rows.append(((2, "C"), _and(_cmp(">=", "hcp", 11), *no_biddable_4, *no_lower_5,
                                    _or(_not(the_fit), _cmp(">=", "hcp", 15)))))
    rows.append(((2, suit), _and(_cmp(">=", "hcp", 6), _cmp("<=", "hcp", 9), the_fit, *no_biddable_4)))
    unspliced = "(" + " ".join(no_lower_5) + ")" if no_lower_5 else "nil"
    rows.append(((3, suit), _and(_cmp(">=", "hcp", 10), _cmp("<=", "hcp", 12), the_fit, *no_biddable_4, unspliced)))
    if o >= 2:
        rows.append(((4, suit), _and(_cmp(">=", "hcp", 13), _cmp("<=", "hcp", 15), the_fit)))
    for i in (1, 2, 3):
        if o > i:
            rows.append(((2, SUITS[i]), _and(_cmp(">=", "hcp", 11), _cmp(">=", SUITS[i], 5))))
    return RuleTable(rows)


# ---- entry points (GPU) ------------------------------------------------------------
def _filler_deal(hand: Hand):
    """a record whose North hand is `hand` (other seats get the remaining cards in order)"""
    rest = [c for c in range(52) if c not in set(hand.cards())]
    others = []
    for k in range(3):
        cards = rest[13 * k:13 * k + 13]
        suits = [[c % 13 for c in cards if c // 13 == s] for s in range(4)]
        others.append(Hand(suits))
    return pack_deal([hand] + others)


def assess_hands(hands, scheme: RuleTable | None = None, engine: Engine | None = None):
    """Batch assess-hand (+ choose-bid over `scheme`) for 13-card hands: structured array[n]."""
    eng = engine or Engine.default()
    rec = np.stack([_filler_deal(h) for h in hands]) if len(hands) else np.empty((0, 2), np.uint64)
    _, feats = eng.filter(rec, None, scheme.scheme() if scheme else None)
    return feats[:, 0]


def assess_hand(hand: Hand, measure=None, engine: Engine | None = None):
    """bidding.cl:33-38: (lens power loosers), or `measure` applied to them."""
    f = assess_hands([hand], None, engine)[0]
    vals = ([int(x) for x in f["len"]], [int(x) for x in f["suit_hcp"]], int(f["losers"]))
    return measure(*vals) if measure else vals


def balanced(hand: Hand, engine: Engine | None = None) -> bool:
    """bidding.cl:40-46 balanced? of a hand"""
    return bool(assess_hands([hand], None, engine)[0]["balanced"])


class BidError(RuntimeError):
    """Evaluation reached a form the reference cannot evaluate (illegal function call)."""


def choose_bid(hand: Hand, scheme: RuleTable = None, engine: Engine | None = None):
    """bidding.cl:146-150: the first bid of `scheme` (default `openings`) whose rule holds, else None."""
    scheme = scheme or openings
    k = int(assess_hands([hand], scheme, engine)[0]["bid"])
    if k == -2:
        raise BidError("illegal function call while evaluating the rule table")
    return None if k < 0 else scheme.rows[k][0]


def test_bid(bid, hand: Hand, scheme: RuleTable = None, engine: Engine | None = None) -> bool:
    """bidding.cl:111-112: does the rule of `bid` (in `openings`) hold for `hand`?"""
    scheme = scheme or openings
    one = RuleTable([(_norm_bid(bid), scheme.form(bid))])
    k = int(assess_hands([hand], one, engine)[0]["bid"])
    if k == -2:
        raise BidError("illegal function call while evaluating the rule")
    return k == 0


# ---- second-round bid generators (bidding.cl:263-413), host-side list manipulation ----------
# Rule forms here are S-expression lists as read by sexp.read (symbols upper-cased); the
# generated tables are turned into RuleTable text for the GPU like every other scheme.
class LispCallError(RuntimeError):
    """The reference would signal a wrong-number-of-arguments error here."""


def _sym(x):
    return sexp.Sym(str(x).upper())


def _form(x):
    return sexp.read(x) if isinstance(x, str) else x


def matchlist(pattern, mapper, lst, args=()):
    """bridge.cl:183-190: match `pattern` against the head of `lst`; `_` captures."""
    if not pattern:
        return mapper(*reversed(args))
    if not lst:
        return None
    if pattern[0] == "_":
        return matchlist(pattern[1:], mapper, lst[1:], (lst[0],) + tuple(args))
    if pattern[0] == lst[0]:
        return matchlist(pattern[1:], mapper, lst[1:], args)
    return None


def match_smaller(sym, lst):
    """bidding.cl:263-265"""
    r = matchlist([_sym("<"), sym, "_"], lambda v: v - 1, lst)
    return r if r is not None else matchlist([_sym("<="), sym, "_"], lambda v: v, lst)


def match_greater(sym, lst):
    """bidding.cl:267-269"""
    r = matchlist([_sym(">"), sym, "_"], lambda v: v + 1, lst)
    return r if r is not None else matchlist([_sym(">="), sym, "_"], lambda v: v, lst)


def match_longer(lst):
    """bidding.cl:271-275.  The `(> _ _)` branch applies a one-argument lambda* to two
    arguments in the reference, i.e. it signals an error; only `(>= suit n)` works."""
    def bad(*_):
        raise LispCallError("match-longer: lambda* applied to 2 arguments (bidding.cl:272)")
    ans = matchlist([_sym(">"), "_", "_"], bad, lst)
    if ans is None:
        ans = matchlist([_sym(">="), "_", "_"], lambda *x: list(x), lst)
    if ans and ans[0] in ("C", "D", "H", "S"):
        return ans
    return None


def and_condition(conds):
    """bidding.cl:277-279"""
    sig = [c for c in conds if c is not None]
    if len(sig) > 1:
        return [_sym("AND")] + sig
    return sig[0] if sig else None


def combine_shapes(a, b):
    """bidding.cl:281-297: add the bounds two partners have shown."""
    a, b = _form(a), _form(b)

    def self(base, other):
        if not base:
            return None
        if base[0] == "AND":
            if other[0] == "AND":
                out = []
                for conj in base[1:]:
                    op, measure, val = (list(conj) + [None, None, None])[:3]
                    extractor = match_greater if op in (">", ">=") else match_smaller
                    otherval = next((v for v in (extractor(measure, o) for o in other[1:]) if v is not None), None)
                    out.append([op, measure, val + otherval] if otherval is not None else None)
                return and_condition(out)
            return self(base, [_sym("AND"), other])
        return self([_sym("AND"), base], other)

    return self(a, b)


def _find_in_bid(meaning, fn):
    """bidding.cl:314-318: only AND/OR meanings work (anything else applies fn to the spread form)"""
    if meaning and meaning[0] in ("AND", "OR"):
        return next((v for v in (fn(m) for m in meaning[1:]) if v is not None), None)
    raise LispCallError("find-in-bid: function applied to a spread form (bidding.cl:318)")


def _filter_bid(meaning, fn):
    """bidding.cl:320-324"""
    if meaning and meaning[0] in ("AND", "OR"):
        return [v for v in (fn(m) for m in meaning[1:]) if v is not None]
    raise LispCallError("filter-bid: function applied to a spread form (bidding.cl:324)")


def max_hcp(meaning):
    return _find_in_bid(_form(meaning), lambda m: match_smaller(_sym("HCP"), m))


def min_hcp(meaning):
    return _find_in_bid(_form(meaning), lambda m: match_greater(_sym("HCP"), m))


def longers(meaning):
    return _filter_bid(_form(meaning), match_longer)


def _suitno_star(suit):
    """bidding.cl:345-348: nil / NT rank above spades"""
    if suit is None or str(suit).upper() == "NT":
        return 4
    return SUITS.index(str(suit).upper())


def game_in(suit):
    """bidding.cl:335-338"""
    if suit is None:
        return (3, "NT")
    return (5, str(suit)) if str(suit) in ("C", "D") else (4, str(suit))


def invite_in(suit):
    """bidding.cl:340-343"""
    if suit is None:
        return (2, "NT")
    return (4, str(suit)) if str(suit) in ("C", "D") else (3, str(suit))


def closest_bid(suit, base):
    """bidding.cl:350-353"""
    base = _norm_bid(base)
    return (base[0], str(suit)) if SUITS.index(str(suit)) > _suitno_star(base[1]) else (base[0] + 1, str(suit))


def jump_bid(suit, base, levels):
    """bidding.cl:355-358"""
    base = _norm_bid(base)
    extra = 0 if SUITS.index(str(suit)) > _suitno_star(base[1]) else 1
    return (base[0] + levels + extra, str(suit))


def bid_jump(base, bid):
    """bidding.cl:360-363"""
    base, bid = _norm_bid(base), _norm_bid(bid)
    return (bid[0] - base[0]) + (0 if _suitno_star(bid[1]) > _suitno_star(base[1]) else -1)


def suit_biddef(bid, lo_hcp, hi_hcp, min_losers, max_losers, length):
    """bidding.cl:371-377 -> (bid, form text)"""
    suit = bid[1]
    return (bid, _and(_or(_cmp(">=", "hcp", lo_hcp), _cmp("<=", "loosers", max_losers)), _cmp("<=", "hcp", hi_hcp),
                      _cmp(">=", "loosers", min_losers), _cmp(">=", suit, 8 - length)))


def further_bid(known_shape, partner_shape, last_bid) -> RuleTable:
    """bidding.cl:379-413: rebids after partner showed `partner_shape`; NIL rows of the
    reference's table (rows that never match) are kept as such."""
    last = _norm_bid(last_bid)
    my_min = min_hcp(known_shape)
    p_min = min_hcp(partner_shape)
    rows = []
    third = p_min // 3
    for suit, length in longers(partner_shape):
        suit = str(suit)
        minor = suit in ("C", "D")
        game_hcp, game_losers = (27, 2) if minor else (25, 3)
        inv_jump = bid_jump(last, invite_in(suit))
        rows.append(suit_biddef(game_in(suit), game_hcp - p_min, 29 - p_min, 2 + third, game_losers + third, length))
        rows.append(suit_biddef(invite_in(suit), game_hcp - p_min - 2, game_hcp - p_min, 3 + third, game_losers + 1 + third,
                                length) if inv_jump > 1 else (None, "nil"))
        rows.append(suit_biddef(jump_bid(suit, last, 0 if inv_jump == 0 else 1), my_min + 3, game_hcp - p_min - 1, 3 + third,
                                game_losers + 1 + third, length) if inv_jump >= 0 else (None, "nil"))
        rows.append(suit_biddef((5, suit), 31 - p_min - 2, 30 - p_min, 1 + third, 1 + third, length))
        if suit != last[1] or last[0] == 1:
            rows.append((closest_bid(suit, last), _cmp(">=", suit, 8 - length)))
    return RuleTable(rows)


class Bidding:
    """bidding.cl:457-480 class `bidding`: finds the first seat (0..3) that has an opening bid,
    rolls the deal so that the opener sits first and selects the response scheme."""

    def __init__(self, deal, engine: Engine | None = None):
        self.deal = list(deal)
        self.bids = None
        self.meanings = None
        self.bid_scheme = openings
        self._engine = engine

    def next(self):
        feats = assess_hands(self.deal, self.bid_scheme, self._engine)
        found = None
        for seat in range(4):
            k = int(feats[seat]["bid"])
            if k == -2:
                raise BidError("illegal function call while evaluating the rule table")
            if k >= 0:
                found = (seat, self.bid_scheme.rows[k][0])
                break
        if not found:
            return None
        passes, bid = found
        from .deal import roll
        self.deal = roll(-passes, self.deal)
        self.bids = [bid]
        self.meanings = [(bid, self.bid_scheme.form(bid)), None]
        if bid == (2, "C"):
            self.bid_scheme = two_c_responses
        elif bid[1] == "NT":
            self.bid_scheme = nt_responses(bid[0])
        elif bid[0] == 1:
            self.bid_scheme = basic_responses(bid)
        else:
            self.bid_scheme = None             # the reference's cond has no live clause here ((nil ...), bidding.cl:479)
        return self.bids


# ---- auctions as generator constraints (BASELINE config "full bidding-sequence constraints") ----
def response_scheme(opening) -> RuleTable | None:
    """The scheme class `bidding` switches to after `opening` (bidding.cl:476-479)."""
    opening = _norm_bid(opening)
    if opening == (2, "C"):
        return two_c_responses
    if opening[1] == "NT":
        return nt_responses(opening[0])
    if opening[0] == 1:
        return basic_responses(opening)
    return None


def auction_rules(dealer: int, calls):
    """Per-seat rule forms (seat order N E S W) for an auction that starts with `dealer` and goes `calls`:
    each call is None / "P" (pass) or a bid like (1, "NT").  Calls before the first bid are judged over
    `openings`; the opener's partner's first call over the response scheme of the opening
    (`response_scheme`); the opponents' calls after the opening are not constrained beyond a pass being a pass
    over `openings` (the reference has no overcall tables).  Returns a list of four forms (None = unconstrained)
    ready for `Constraints(None, rules)`.

    Example: auction_rules(0, ["P", "P", (1, "NT"), "P", (3, "NT")]) -- North deals and passes, East
    passes, South opens 1NT, West passes, North answers 3NT."""
    forms = [[] for _ in range(4)]
    opener = opening = None
    responded = False
    for i, call in enumerate(calls):
        seat = (dealer + i) % 4
        is_pass = call is None or (isinstance(call, str) and call.upper() in ("P", "PASS"))
        if opener is None:
            if is_pass:
                forms[seat].append(openings.passes())
            else:
                opener, opening = seat, _norm_bid(call)
                forms[seat].append(openings.chooses(opening))
        elif seat == (opener + 2) % 4 and not responded:
            responded = True
            scheme = response_scheme(opening)
            if scheme is None:
                raise ValueError(f"no response scheme after {opening}")
            forms[seat].append(scheme.passes() if is_pass else scheme.chooses(call))
        elif is_pass:
            if seat != opener and not forms[seat]:
                forms[seat].append(openings.passes())        # an opponent who has not spoken yet: no opening bid either
        else:
            raise ValueError("only the opening and the first response are modelled by the reference's tables")
    return [None if not fs else fs[0] if len(fs) == 1 else _and(*fs) for fs in forms]


def first_two_calls(records, engine: Engine | None = None, dealer: int = 0):
    """Batched class-`bidding` `next` (bidding.cl:463-480) over deal records: for every deal the first seat, from the
    dealer on, that has an opening bid, that bid, and partner's answer over the response scheme the opening selects.
    One library call (bmc_auction: the whole driver runs in one kernel).
    Returns (opener int8[n] (-1: passed out), opening list, response list)."""
    eng = engine or Engine.default()
    tables = [response_scheme(b) for b in openings.bids()]
    out = eng.auction(records, openings.scheme(), [t.scheme() if t is not None else None for t in tables], dealer)
    if (out["status"] == 1).any():
        raise BidError("illegal function call while evaluating the openings")
    names = openings.bids()
    opening = [None if k < 0 else names[k] for k in out["opening"].tolist()]
    response = []
    for k, r, st in zip(out["opening"].tolist(), out["response"].tolist(), out["status"].tolist()):
        if st == 2:
            response.append("error")
        elif k < 0 or r < 0 or tables[k] is None:
            response.append(None)
        else:
            response.append(tables[k].rows[r][0])
    return out["opener"].copy(), opening, response
