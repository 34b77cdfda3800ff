"""Host-side logic on CPU: S-expression reader, cards, hist-* pipeline, infer-*,
contract parsing, the rule tables, and the C-ABI library's exports (no compute
calls: there is no GPU here and the library has no CPU fallback)."""
import ctypes as C
import os
import re
from fractions import Fraction

import numpy as np
import pytest

from kat_util import bid_of, hand_of, load_kats, truth, unquote
from bridge_mc_b200 import _lib, bidding, cards, hist, infer, sexp
from bridge_mc_b200.compscore import Contract, scanstr
from bridge_mc_b200.deal import Deal, close_list, invert_id, normdef, reorder, roll
from bridge_mc_b200.engine import Constraints, Scheme, seat_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def norm(x):
    """nil == (): normalise None/[] and Sym/str for comparisons with KAT data"""
    if x is None or x == []:
        return None
    if isinstance(x, list):
        if len(x) == 2 and x[0] == "QUOTE":          # ''foo inside quoted data stays (quote foo)
            return ["QUOTE", norm(x[1])]
        return [norm(e) for e in x]
    return x


# ---- sexp / cards -----------------------------------------------------------------
def test_sexp_roundtrip():
    txt = "(((2 C) . (or (> hcp 22) (and (not balanced) (<= loosers 4)))) ((1 NT) . balanced) (a . b) \"x y\" 3/4 nil)"
    v = sexp.read(txt)
    assert v[0][0] == [2, "C"] and v[0][1] == "OR"
    assert isinstance(v[1], sexp.Dotted) and v[1].tail == "BALANCED"
    assert v[4] == Fraction(3, 4) and v[5] is None
    assert sexp.read(sexp.dumps(v)) == v
    assert sexp.read("( :hcp ( 15 17 ) :S ( 5 NIL ( 3 NIL ) ) )") == [":HCP", [15, 17], ":S", [5, None, [3, None]]]


def test_str2hand_kats():
    for where, form, exp in load_kats("SUITS"):
        assert cards.str2hand(str(form[1][1])).suits == exp, where
    for where, form, exp in load_kats("HANDID"):
        assert cards.str2hand(str(form[1][1])).id == str(exp), where
    for where, form, exp in load_kats("MAPCAR"):
        deal = cards.str2deal(str(form[2][1]))
        if form[1] == ["FUNCTION", "SUITS"]:
            assert [h.suits for h in deal] == exp, where
        else:
            assert [h.id for h in deal] == [str(x) for x in exp], where
    for where, form, exp in load_kats("CARD"):
        assert list(cards.card(str(form[1]))) == exp, where
    h = cards.str2hand("♣ 72 ♦ KJ862 ♥  ♠ AQ8652")
    assert h.suits[2] == [] and len(h) == 13
    assert cards.print_hand(h.suits) == "♣ 72 ♦ KJ862 ♥  ♠ AQ8652"


def test_record_pack_unpack():
    deal = cards.str2deal("N: ♣ KJ96 ♦ A8 ♥ AJ83 ♠ AK9\nE: ♣ Q5 ♦ J9762 ♥ Q ♠ J6543\n"
                          "S: ♣ 432 ♦ 1043 ♥ K10954 ♠ 87\nW: ♣ A1087 ♦ KQ5 ♥ 762 ♠ Q102")
    rec = cards.pack_deal(deal)
    back = cards.unpack_deal(rec)
    assert [h.suits for h in back] == [h.suits for h in deal]
    with pytest.raises(ValueError):
        cards.pack_deal([deal[0], deal[0], deal[2], deal[3]])


def test_list_helpers_kats():
    for where, form, exp in load_kats("INVERT-ID"):
        assert invert_id(unquote(form[1])) == exp, where
    for where, form, exp in load_kats("REORDER"):
        assert reorder(unquote(form[1]), unquote(form[2])) == exp, where
    assert roll(2, [10, 100, 15, 5]) == [15, 5, 10, 100] and roll(-1, [10, 100, 15, 5]) == [100, 15, 5, 10]
    assert close_list([2, 0], [0, 1, 2, 3]) == [2, 0, 2, 3]
    assert normdef({"hcp": 15, "s": [5, None, 3]}) == {"hcp": [15, 15], "s": [5, None, [3, 3]]}


# ---- hist-* -------------------------------------------------------------------------
def test_hist_kats():
    n = 0
    for where, form, exp in load_kats("HIST-BASE"):
        args = [unquote(a) for a in form[1:]]
        assert norm(hist.hist_base(*args)) == norm(exp), where
        n += 1
    for where, form, exp in load_kats("HIST-BREAKDOWN"):
        assert norm(hist.hist_breakdown(unquote(form[1]))) == norm(exp), where
        n += 1
    for where, form, exp in load_kats("HIST-PERCENT"):
        assert norm(hist.hist_percent(unquote(form[1]))) == norm(exp), where
        n += 1
    for where, form, exp in load_kats("HISTOGRAM"):
        assert norm(hist.histogram(unquote(form[1]))) == norm(exp), where
        n += 1
    assert n == 3 + 2 + 1 + 1


def test_hist_base_accepts_device_precounts():
    base = hist.tallies_as_base([0, 5, 0, 2], label=lambda i: 10 + i)
    assert base == [[11, 5], [13, 2]]
    assert hist.hist_base([13, 14], base) == [[11, 5], [13, 3], [14, 1]]
    lines = hist.print_hist(hist.hist_percent([[1, 999], [2, 1], [3, 1000000]]), out=[])
    assert lines[0].startswith("3: 99.90%") and lines[-1] == "2: 1/1001000"


# ---- infer-* --------------------------------------------------------------------------
def _plist(x):
    return infer.plist_to_def(unquote(x))


def _expect_def(exp):
    return infer.plist_to_def(exp) if exp else {}


def test_infer_kats():
    n = 0
    full = [[9, 1, 1, 1, 1]] * 4
    for head, fn in (("INFER-HCP-RANGE", infer.infer_hcp_range), ("INFER-SUIT-RANGES", infer.infer_suit_ranges)):
        for where, form, exp in load_kats(head):
            sup = form[1]
            sup = full if sup == ["SUPPLY", ["ALL-CARDS"]] else unquote(sup)
            got = fn(sup, _plist(form[2]), *[_plist(o) for o in form[3:]])
            assert norm({k: v for k, v in got.items()}) == norm(_expect_def(exp)) or got == _expect_def(exp), where
            assert list(got.keys()) == list(_expect_def(exp).keys()), where
            n += 1
    for where, form, exp in load_kats("INFER-DISTGEN-PARAMS"):
        got = infer.infer_distgen_params(unquote(form[1]), _plist(form[2]), [_plist(o) for o in (unquote(form[3]) or [])])
        assert got == _expect_def(exp) and list(got) == list(_expect_def(exp)), where
        n += 1
    assert n == 7 + 5 + 2


def test_infer_bad_range():
    full = [[9, 1, 1, 1, 1]] * 4
    with pytest.raises(infer.BadRangeError):
        infer.infer_hcp_range(full, {"hcp": [30, 35]}, {"hcp": [12, 15]})
    with pytest.raises(_lib.BadRange):
        infer.infer_suit_ranges(full, {"s": [10, None]}, {"s": [5, 6]})
    with pytest.raises(infer.BadRangeError):
        Deal(defs=[{"hcp": [30, 37]}, {"hcp": [15, 17]}]).constraints()


# ---- compscore parsing --------------------------------------------------------------------
def test_scanstr_and_contract_kats():
    for where, form, exp in load_kats("SCANSTR"):
        opts = [str(o) if not isinstance(o, int) else o for o in unquote(form[1])]
        got, rest = scanstr(opts, str(form[2]))
        assert [got, rest] == [exp[0] if isinstance(exp[0], int) else str(exp[0]), str(exp[1])], where
    for where, form, exp in load_kats("FORMAT", file="compscore.cl"):
        assert str(Contract(str(form[3][3]))) == str(exp), where
    c = Contract("4HSx+1")
    assert (c.level, c.suit, c.declarer, c.double, c.outcome, c.tricks()) == (4, "H", "S", "X", "+1", 11)
    assert Contract("3NT").outcome == "=" and Contract("3NT").declarer is None
    assert Contract("6♥Wx-2").tricks() == 10
    assert Contract("3NTS").with_score(10).outcome == 1


# ---- rule tables ------------------------------------------------------------------------------
def test_rule_tables_compile_and_shape():
    assert [b for b in bidding.openings.bids()][:3] == [(2, "C"), (2, "NT"), (1, "NT")]
    assert len(bidding.openings.rows) == 14
    assert bidding.nt_responses(1).bids() == [(2, "C"), (2, "D"), (2, "H"), (2, "S"), (3, "C"), (2, "NT"), (3, "NT"),
                                              (4, "NT"), (6, "NT")]
    assert bidding.nt_responses(2).bids() == [(3, "C"), (3, "D"), (3, "H"), (3, "S"), (3, "NT"), (4, "NT"), (6, "NT")]
    assert bidding.basic_responses((1, "H")).bids() == [(1, "S"), (1, "NT"), (2, "C"), (2, "H"), (3, "H"), (4, "H"),
                                                        (2, "D")]
    assert bidding.basic_responses("(1 C)").bids()[:3] == [(1, "D"), (1, "H"), (1, "S")]
    with pytest.raises(ValueError):
        bidding.basic_responses((2, "H"))
    for t in (bidding.openings, bidding.nt_responses(1), bidding.nt_responses(2), bidding.two_c_responses,
              bidding.basic_responses((1, "S")), bidding.basic_responses((1, "C"))):
        assert t.scheme().size == len(t.rows)
        assert sexp.read(t.sexpr())          # well-formed Lisp
    assert "(not (or" in bidding.openings.chooses((1, "NT")) and bidding.openings.passes().startswith("(not (or")


# ---- the C-ABI library -------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bridgemc.h")).read()
    declared = sorted(set(re.findall(r"\b(bmc_[a-z_0-9]+)\s*\(", hdr)))
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), f"libbridgemc.so does not export {name}"
    assert sorted(_lib.EXPORTS) == declared
    assert C.sizeof(_lib.SeatSpec) == 32 and C.sizeof(_lib.Stats) == 48 and C.sizeof(_lib.Contract) == 4
    assert _lib.TALLY_WORDS == 4 + 4 * 40 + 4 * 4 * 14 + 4 * 40 + 8 * 14 + 8 * 49 + 8
    assert b"sm_100a" in L.bmc_version()


def test_shape_table():
    t = _lib.shape_table()
    assert t.shape == (39, 4) and (t.sum(axis=1) == 13).all()
    assert all(r[0] >= r[1] >= r[2] >= r[3] for r in t.tolist())
    assert len({tuple(r) for r in t.tolist()}) == 39 and tuple(t[0]) == (4, 3, 3, 3)


def test_constraint_compile_errors():
    Constraints([seat_spec(hcp=(15, 17), s=(5, None, (3, None))), None, None, None])
    Constraints(None, [None, None, bidding.openings.chooses((1, "NT")), bidding.openings.passes()])
    with pytest.raises(_lib.BadRange):
        Constraints([seat_spec(hcp=(20, 10))])
    with pytest.raises(_lib.BadRange):
        Constraints([seat_spec(hcp=(0, 5))], ["(> hcp 22)"])
    with pytest.raises(_lib.BadRange):
        Constraints([seat_spec(hcp=(15, None)), seat_spec(hcp=(15, None)), seat_spec(hcp=(15, None))])
    with pytest.raises(_lib.BridgeMcError) as e:
        Constraints(None, ["(launch-missiles)"])
    assert e.value.code == _lib.BMC_E_PARSE
    with pytest.raises(_lib.BridgeMcError):
        Scheme("((1 NT) . ")
    h = cards.str2hand("♣ AJ7 ♦ AK54 ♥ 6543 ♠ 64")
    Constraints([seat_spec(fixed=h)])
    with pytest.raises(_lib.BridgeMcError):
        Constraints([seat_spec(fixed=h), seat_spec(fixed=h)])


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bridge_mc_b200.engine import Engine
    with pytest.raises(_lib.BridgeMcError) as e:
        Engine(0)
    assert e.value.code == _lib.BMC_E_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bridge_mc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".sh")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in text.lower().replace("oracle/bmc_oracle.c", "").replace("orc_gpu_deal", "") \
                    or f in ("deal.cuh",), f


# ---- second-round generators (bidding.cl:263-413) -------------------------------------------
def test_bid_jump_and_combine_shapes_kats():
    n = 0
    for where, form, exp in load_kats("BID-JUMP"):
        assert bidding.bid_jump(bid_of(form[1]), bid_of(form[2])) == exp, where
        n += 1
    for where, form, exp in load_kats("COMBINE-SHAPES"):
        assert bidding.combine_shapes(unquote(form[1]), unquote(form[2])) == exp, where
        n += 1
    assert n == 5 + 3
    assert bidding.min_hcp("(and (>= hcp 12) (<= hcp 22) (>= D 5))") == 12
    assert bidding.max_hcp("(and (>= hcp 12) (< hcp 23))") == 22
    assert bidding.longers("(and (>= hcp 6) (>= S 4) (>= H 5))") == [["S", 4], ["H", 5]]
    with pytest.raises(bidding.LispCallError):
        bidding.min_hcp("(>= hcp 6)")                    # bidding.cl:318 applies fn to the spread form
    with pytest.raises(bidding.LispCallError):
        bidding.longers("(and (> S 4))")                 # bidding.cl:272 lambda* arity
    assert bidding.game_in(None) == (3, "NT") and bidding.game_in("D") == (5, "D") and bidding.invite_in("S") == (3, "S")
    assert bidding.closest_bid("S", (1, "H")) == (1, "S") and bidding.closest_bid("C", (1, "H")) == (2, "C")
    assert bidding.jump_bid("S", (1, "S"), 1) == (3, "S")


def test_further_bid_table_shape():
    t = bidding.further_bid("(and (>= hcp 12) (<= hcp 22) (>= D 5))", "(and (>= hcp 6) (>= S 4))", (1, "S"))
    assert t.bids() == [(4, "S"), None, (3, "S"), (5, "S"), (2, "S")]        # NIL row kept: invite-jump = 1
    t2 = bidding.further_bid("(and (>= hcp 6) (>= S 4))", "(and (>= hcp 12) (<= hcp 22) (>= D 5) (>= S 4))", (2, "S"))
    assert t2.bids() == [(5, "D"), None, (4, "D"), (5, "D"), (3, "D"), (4, "S"), None, (3, "S"), (5, "S")]
    assert t2.scheme().size == len(t2.rows)


def test_unranking_arithmetic_selftest(tmp_path):
    """tests/unrank_selftest.cu: the seat-first sampler's unranking helpers (seatfirst.cuh) compiled for the HOST --
    exact division by S and by the class multiplicities, bijectivity of the honour and low-card maps, the
    compile-time binomial tables"""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "unrank_selftest")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unrank_selftest.cu")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "unrank_selftest: ok" in out.stdout, out.stdout[-2000:]


def test_primary_seat_and_its_hcp_window_without_a_gpu():
    """bmc_constraints_create is host code: the seat the seat-first sampler draws by rank, and the HCP window its
    class table is ordered by (ranges of the spec, tightened by what the rule form implies)"""
    from bridge_mc_b200.engine import Constraints, seat_spec
    c = Constraints(None, [None, None, bidding.openings.form((1, "NT")), None])
    assert (c.primary, c.primary_hcp, c.mode) == (2, (15, 17), 0)
    c = Constraints([None, seat_spec(hcp=(20, 21)), None, seat_spec(hcp=(0, 30))])
    assert (c.primary, c.primary_hcp) == (1, (20, 21))
    c = Constraints([seat_spec(s=(6, 13)), None, None, None])
    assert (c.primary, c.primary_hcp) == (0, (0, 37))
    c = Constraints([None, None, None, seat_spec(hcp=(12, None))], [None, None, None, "(< hcp 15)"])
    assert (c.primary, c.primary_hcp) == (3, (12, 14))


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "sample_1nt")
    libdir = os.path.join(root, "bridge_mc_b200")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "sample_1nt.c"), "-L" + libdir, "-lbridgemc", "-Wl,-rpath," + libdir,
                    "-o", exe], check=True)
    return exe


def test_c_abi_from_plain_c_without_a_gpu(tmp_path):
    """include/bridgemc.h is C99 and examples/sample_1nt.c links against the library; without a device the
    program reports BMC_E_CUDA (exit 3) -- no CPU fallback"""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the gpu-marked variant")
    out = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 3 and "needs a CUDA device" in out.stdout, out.stdout


@pytest.mark.gpu
def test_c_abi_from_plain_c(tmp_path):
    import subprocess
    out = subprocess.run([_build_c_example(tmp_path), "50000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mode 4, primary seat 2, HCP window 15-17" in out.stdout      # auto: exact seat-first
    m = re.search(r"accepted (\d+)\s+stored (\d+)", out.stdout)
    assert m and int(m.group(1)) >= 50000 and int(m.group(2)) == 50000
    assert out.stdout.count("\n") >= 7 and "S:" in out.stdout


def test_rest_parsers():
    """rest.cl:279-303: hand-or-def-parser and the shape of measure-parser's result (the trick estimator behind
    msr-<contract> runs on the GPU, tests/test_gpu_playdeal.py; here a stand-in is passed)"""
    from bridge_mc_b200.deal import hand_or_def_parser, measure_parser
    south = "S: ♠ AKJ5 ♥ K1064 ♦ Q87 ♣ A6"
    defs = hand_or_def_parser([(":|hand|", south), ("def", "(:hcp (10 12) :s (4 nil))"), ("def", "NIL"), ("def", "(:hcp 12)")])
    assert defs[0] == cards.str2hand(south)
    assert defs[1:] == [{"hcp": [10, 12], "s": [4, None]}, {}, {"hcp": 12}]
    assert [m.__name__ for m in measure_parser([""])] == ["best_outcomes"]       # the default measure (rest.cl:288)
    ms = measure_parser(["3NTS", "4HNx"], best_tricks=lambda suit, a, b, c, d: 9)
    assert [m.__name__ for m in ms] == ["msr-3NTS", "msr-4HNx"]


@pytest.mark.gpu
def test_measure_parser_scores_through_the_library():
    from bridge_mc_b200.deal import Deal, McCase, measure_parser
    from bridge_mc_b200.engine import Engine
    eng = Engine.default()
    seen = []

    def best_tricks(suit, declarer, left, dummy, right):
        seen.append((suit, declarer))
        return 10 if suit == "H" else 9

    ms = measure_parser(["3NTS", "4HNx"], best_tricks=best_tricks, engine=eng)
    mc = McCase(Deal(engine=eng, seed=11), ms)
    rows = mc.sim(3)
    for hands, a, b in rows:
        assert a == [9, 400] and b == [10, 590]            # 3NT= and 4Hx= not vulnerable (compscore.cl KAT values)
    # the declarer's hand comes first: South for 3NTS, North for 4HNx
    assert seen[0] == (None, rows[0][0][2]) and seen[1] == ("H", rows[0][0][0])


def test_deal_path_list_utilities_kats():
    """the reference's inline tests of the small list utilities its deal path is written with
    (bridge_mc_b200/lists.py): take, remove-cards, cards-by, valid-hcp-sums, push-pos, roll, zip-deals,
    even-sublists, suit<>, card<>; mod-or-push by hand (its test forms hold lambdas)"""
    import operator
    from bridge_mc_b200 import lists
    fns = {"CARD-HCP": lists.card_hcp, "CARD-SUITNO": lists.card_suitno, "<": operator.lt, ">": operator.gt,
           "=": operator.eq, "<=": operator.le, ">=": operator.ge}

    def ev(x):
        if isinstance(x, list) and x and x[0] == "QUOTE":
            return x[1]
        if isinstance(x, list) and x and x[0] == "FUNCTION":
            return fns[str(x[1])]
        if x == ["ALL-CARDS"]:
            return lists.all_cards()
        return x

    def kwargs(rest):
        out = {}
        for i in range(0, len(rest), 2):
            out[str(rest[i]).lstrip(":").lower()] = ev(rest[i + 1])
        return out

    def nil(x):                       # NIL and () are the same object in Lisp
        if x is None:
            return []
        return [nil(y) for y in x] if isinstance(x, list) else x

    seen = 0
    for where, form, exp in load_kats("TAKE"):
        assert lists.take(form[1], ev(form[2])) == exp, where
        seen += 1
    for where, form, exp in load_kats("REMOVE-CARDS"):
        assert lists.remove_cards(ev(form[1]), ev(form[2])) == exp, where
        seen += 1
    for where, form, exp in load_kats("CARDS-BY"):
        assert nil(lists.cards_by(ev(form[1]), ev(form[2]))) == nil(exp), where
        seen += 1
    for where, form, exp in load_kats("VALID-HCP-SUMS"):
        assert lists.valid_hcp_sums(form[1], **kwargs(form[2:])) == exp, where
        seen += 1
    for where, form, exp in load_kats("PUSH-POS"):
        assert nil(lists.push_pos(ev(form[1]), form[2], ev(form[3]))) == nil(exp), where
        seen += 1
    for where, form, exp in load_kats("ROLL"):
        assert lists.roll(form[1], ev(form[2])) == exp and roll(form[1], ev(form[2])) == exp, where
        seen += 1
    for where, form, exp in load_kats("ZIP-DEALS"):
        assert lists.zip_deals(ev(form[1])) == exp, where
        seen += 1
    for where, form, exp in load_kats("EVEN-SUBLISTS"):
        assert lists.even_sublists(form[1], ev(form[2])) == exp, where
        seen += 1
    for head in ("<", ">", "="):
        for where, form, exp in load_kats(head):
            call = form[1]
            f = lists.suit_cmp if call[0] == "SUIT<>" else lists.card_cmp
            assert fns[head](f(ev(call[1]), ev(call[2])), form[2]) is truth(exp), where
            seen += 1
    assert seen == 3 + 1 + 3 + 7 + 4 + 4 + 1 + 1 + 7
    # bridge.cl:466-484
    dbl = lambda ref, x: x * 2 if ref == x else None
    assert lists.mod_or_push(dbl, lambda x: x, [1, 4, 5, 6, 10], 1) == [2, 4, 5, 6, 10]
    assert lists.mod_or_push(dbl, lambda x: x, [1, 4, 5, 6, 10], 7) == [1, 4, 5, 6, 10, 7]
    assert lists.mod_or_push(dbl, lambda x: [x, 1], [], 1) == [[1, 1]]
    with pytest.raises(lists.BadIndex):
        lists.take(3, [1, 2, 3])
    # bridge.cl:250-258 divlist, and supply (876-878) against the bit-plane version the library's callers use
    lt10, gt5 = (lambda x: 10 < x), (lambda x: 5 > x)
    assert lists.divlist([lt10, gt5], [5, 10, 15, 20, 11, 1, 0, 3, 17, 20]) == [[20, 17, 11, 20, 15], [3, 0, 1], [10, 5]]
    assert lists.divlist([lt10, gt5], [5, 10, 15, 20, 11, 1, 0, 3, 17, 20], map=lambda x: 2 * x) == [[40, 34, 22, 40, 30], [6, 0, 2], [20, 10]]
    some = [c for i, c in enumerate(lists.all_cards()) if i % 3 != 1]
    assert lists.supply(some) == infer.supply([13 * lists.suitno(s) + r for s, r in some])
    assert lists.supply(lists.all_cards()) == [[9, 1, 1, 1, 1]] * 4
