"""ctypes front-end of the CPU oracle (oracle/bmc_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of bmc_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by the
product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libbmc_oracle.so")
    src = os.path.join(HERE, "bmc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return so


class Spec(C.Structure):
    """orc_spec: -1 = nil."""
    _fields_ = [("hcp_lo", C.c_int), ("hcp_hi", C.c_int), ("has_hcp", C.c_int),
                ("len_lo", C.c_int * 4), ("len_hi", C.c_int * 4),
                ("shcp_lo", C.c_int * 4), ("shcp_hi", C.c_int * 4), ("has_shcp", C.c_int * 4)]


def make_spec(hcp=None, c=None, d=None, h=None, s=None) -> Spec:
    """Range spec in the reference's keyword form: number | (lo hi) | (lo hi (slo shi))."""
    sp = Spec()
    sp.has_hcp = 0
    sp.hcp_lo = sp.hcp_hi = -1
    if hcp is not None:
        sp.has_hcp = 1
        lo, hi = (hcp, hcp) if isinstance(hcp, int) else (hcp[0], hcp[1])
        sp.hcp_lo = -1 if lo is None else lo
        sp.hcp_hi = -1 if hi is None else hi
    for i, spec in enumerate((c, d, h, s)):
        sp.len_lo[i] = sp.len_hi[i] = -1
        sp.shcp_lo[i] = sp.shcp_hi[i] = -1
        sp.has_shcp[i] = 0
        if spec is None:
            continue
        if isinstance(spec, int):
            sp.len_lo[i] = sp.len_hi[i] = spec
            continue
        sp.len_lo[i] = -1 if spec[0] is None else spec[0]
        sp.len_hi[i] = -1 if spec[1] is None else spec[1]
        if len(spec) > 2 and spec[2] is not None:
            sp.has_shcp[i] = 1
            sh = spec[2]
            lo, hi = (sh, sh) if isinstance(sh, int) else (sh[0], sh[1])
            sp.shcp_lo[i] = -1 if lo is None else lo
            sp.shcp_hi[i] = -1 if hi is None else hi
    return sp


FULL_DECK = (1 << 52) - 1


class Def(C.Structure):
    """orc_def: a range def as `build` sees it (numbers normalised to (n n), like rest.cl's normdef); -1 = nil."""
    _fields_ = [("has_hcp", C.c_int), ("hcp_lo", C.c_int), ("hcp_hi", C.c_int),
                ("has_suit", C.c_int * 4), ("len_lo", C.c_int * 4), ("len_hi", C.c_int * 4),
                ("has_third", C.c_int * 4), ("th_lo", C.c_int * 4), ("th_hi", C.c_int * 4)]


def _pair(v):
    if isinstance(v, int):
        return v, v
    lo = -1 if v[0] is None else int(v[0])
    hi = -1 if len(v) < 2 or v[1] is None else int(v[1])
    return lo, hi


def make_def(d) -> Def:
    """{'hcp': [lo, hi] | n, 'c': [lo, hi, [slo, shi]] | n, ...} (None = nil) -> Def"""
    out = Def()
    out.hcp_lo = out.hcp_hi = -1
    for i in range(4):
        out.len_lo[i] = out.len_hi[i] = out.th_lo[i] = out.th_hi[i] = -1
    d = d or {}
    if d.get("hcp") is not None:
        out.has_hcp = 1
        out.hcp_lo, out.hcp_hi = _pair(d["hcp"])
    for i, k in enumerate("cdhs"):
        v = d.get(k)
        if v is None:
            continue
        out.has_suit[i] = 1
        out.len_lo[i], out.len_hi[i] = _pair(v)
        if not isinstance(v, int) and len(v) > 2 and v[2] is not None:
            out.has_third[i] = 1
            out.th_lo[i], out.th_hi[i] = _pair(v[2])
    return out


def def_to_dict(d: Def) -> dict:
    nil = lambda x: None if x < 0 else x
    out = {}
    if d.has_hcp:
        out["hcp"] = [nil(d.hcp_lo), nil(d.hcp_hi)]
    for i, k in enumerate("cdhs"):
        if d.has_suit[i]:
            v = [nil(d.len_lo[i]), nil(d.len_hi[i])]
            if d.has_third[i]:
                v.append([nil(d.th_lo[i]), nil(d.th_hi[i])])
            out[k] = v
    return out


class Contract(C.Structure):
    _fields_ = [("level", C.c_int8), ("suit", C.c_int8), ("declarer", C.c_int8), ("dbl", C.c_int8)]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    u16x4 = C.c_uint16 * 4
    i11 = C.c_int * 11
    L.orc_suit_hcp.argtypes = [C.c_uint]
    L.orc_suit_losers.argtypes = [C.c_uint]
    L.orc_assess.argtypes = [u16x4, i11]
    L.orc_assess.restype = None
    L.orc_choose_bid.argtypes = [C.c_int, u16x4]
    L.orc_test_bid.argtypes = [C.c_int, C.c_int, u16x4]
    L.orc_test_rule.argtypes = [C.c_int, C.c_int, i11]
    L.orc_scheme_size.argtypes = [C.c_int]
    L.orc_contract_score.argtypes = [C.c_int] * 5
    L.orc_base_score.argtypes = [C.c_int] * 6
    L.orc_imps.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.orc_combinations.argtypes = [C.c_int, C.c_int]
    L.orc_combinations.restype = C.c_uint64
    L.orc_supply.argtypes = [C.c_uint64, (C.c_int * 5) * 4]
    L.orc_supply.restype = None
    L.orc_distgen_new.argtypes = [C.c_uint64, C.POINTER(Spec)]
    L.orc_distgen_new.restype = C.c_void_p
    L.orc_distgen_free.argtypes = [C.c_void_p]
    L.orc_distgen_free.restype = None
    L.orc_distgen_npat.argtypes = [C.c_void_p]
    L.orc_distgen_total.argtypes = [C.c_void_p]
    L.orc_distgen_total.restype = C.c_uint64
    L.orc_distgen_item.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint)]
    L.orc_distgen_item.restype = C.c_uint64
    L.orc_distgen_pmf.argtypes = [C.c_void_p, C.c_uint64 * 38, (C.c_uint64 * 14) * 4]
    L.orc_distgen_pmf.restype = None
    L.orc_rand_hand.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.orc_rand_hand.restype = C.c_uint64
    L.orc_deal_all.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.orc_build_one.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64 * 4]
    L.orc_build_one.restype = None
    L.orc_philox4x32_10.argtypes = [C.c_uint32 * 4, C.c_uint32 * 2, C.c_uint32 * 4]
    L.orc_philox4x32_10.restype = None
    L.orc_gpu_deal.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.orc_features_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_features_batch.restype = None
    L.orc_choose_bid_batch.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_choose_bid_batch.restype = None
    L.orc_gpu_deal_batch.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    L.orc_gpu_deal_batch.restype = None
    L.orc_gpu_deal_fixed_batch.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64 * 4, C.c_void_p, C.c_void_p]
    L.orc_gpu_deal_fixed_batch.restype = None
    L.orc_gpu_deal_seatfirst_batch.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_gpu_deal_seatfirst_batch.restype = None
    L.orc_cpu_best_run.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    L.orc_cpu_best_run.restype = C.c_uint64
    L.orc_ref_run.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.orc_ref_run.restype = C.c_uint64
    L.orc_build_batch.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64, C.c_void_p]
    L.orc_build_batch.restype = None
    L.orc_tricks_of.argtypes = [C.c_uint32 * 2, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32 * 8]
    L.orc_tricks_of.restype = None
    L.orc_score_tally_records.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(Contract), C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_int, C.c_void_p]
    L.orc_score_tally_records.restype = None
    L.orc_score_tally_index.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Contract), C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_int, C.c_void_p]
    L.orc_score_tally_index.restype = None
    L.orc_infer_distgen_params.argtypes = [(C.c_int * 5) * 4, C.POINTER(Def), C.POINTER(Def), C.c_int, C.POINTER(Def)]
    L.orc_ref_build_batch.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.POINTER(Def), C.c_int, C.c_uint64, C.c_uint64, C.c_int,
                                      C.c_void_p, C.c_void_p]
    L.orc_xtab_build.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int * 11, C.c_int * 11]
    L.orc_xtab_build.restype = C.c_void_p
    L.orc_xtab_free.argtypes = [C.c_void_p]
    L.orc_xtab_free.restype = None
    L.orc_xtab_weight.argtypes = [C.c_void_p]
    L.orc_xtab_weight.restype = C.c_uint64
    L.orc_xtab_classes.argtypes = [C.c_void_p]
    L.orc_xtab_classes.restype = C.c_uint32
    L.orc_gpu_deal_exact_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_gpu_deal_exact_batch.restype = None
    _LIB = L
    return L


def infer_distgen_params(supply, base, others):
    """bridge.cl:1267-1269 through the C restatement: supply = 4 x [n_low nJ nQ nK nA] (fewer rows are padded with
    empty suits); defs as make_def takes them.  Returns the narrowed def as a dict, or raises ValueError (bad-range)."""
    sup = ((C.c_int * 5) * 4)()
    for i, row in enumerate(supply):
        for j, v in enumerate(row):
            sup[i][j] = v
    oth = (Def * max(1, len(others)))(*[make_def(o) for o in others])
    out = Def()
    b = make_def(base) if base is not None else None
    rc = lib().orc_infer_distgen_params(sup, C.byref(b) if b is not None else None, oth, len(others), C.byref(out))
    if rc:
        raise ValueError("bad-range")
    return def_to_dict(out)


def ref_build_batch(consts, defs, n: int, seed: int = 1, threads: int = 0):
    """n x `build` (bridge.cl:1292-1319) with the reference's own algorithm: consts = 52-bit id-order masks of the :=
    hands, defs = the :~ defs.  Returns (hands uint64[n,4] in build's order, count int8[n]: hands returned, -1 = error)."""
    import os
    if threads <= 0:
        threads = max(1, len(os.sched_getaffinity(0)))
    carr = (C.c_uint64 * max(1, len(consts)))(*[int(c) for c in consts])
    darr = (Def * max(1, len(defs)))(*[make_def(d) for d in defs])
    out = np.zeros((n, 4), dtype=np.uint64)
    cnt = np.zeros(n, dtype=np.int8)
    lib().orc_ref_build_batch(carr, len(consts), darr, len(defs), n, seed, threads, out.ctypes.data, cnt.ctypes.data)
    return out, cnt


def _score_args(contracts, vulnerable, pairs):
    carr = (Contract * len(contracts))(*[Contract(int(c.level), int(c.suit), int(c.declarer), int(c.dbl)) for c in contracts])
    vul = np.ascontiguousarray(vulnerable, dtype=np.uint8)
    pr = np.ascontiguousarray(pairs, dtype=np.uint8).reshape(-1)
    return carr, vul, pr


def score_tally_records(seed: int, records: np.ndarray, contracts, vulnerable, pairs):
    """CPU mirror of BMC_TALLY_SCORE: (tricks int64[8,14], imps int64[8,49], dropped int64[8]) over the given records."""
    rec = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, 2)
    carr, vul, pr = _score_args(contracts, vulnerable, pairs)
    hist = np.zeros(8 * 14 + 8 * 49 + 8, dtype=np.uint64)
    lib().orc_score_tally_records(seed, rec.ctypes.data, rec.shape[0], carr, vul.ctypes.data, len(contracts), pr.ctypes.data,
                                  pr.size // 2, hist.ctypes.data)
    h = hist.astype(np.int64)
    return h[:112].reshape(8, 14), h[112:112 + 392].reshape(8, 49), h[504:]


def score_tally_index(seed: int, first: int, n: int, contracts, vulnerable, pairs):
    """CPU mirror of bmc_score_tally (index-keyed trick source)."""
    carr, vul, pr = _score_args(contracts, vulnerable, pairs)
    hist = np.zeros(8 * 14 + 8 * 49 + 8, dtype=np.uint64)
    lib().orc_score_tally_index(seed, first, n, carr, vul.ctypes.data, len(contracts), pr.ctypes.data, pr.size // 2,
                                hist.ctypes.data)
    h = hist.astype(np.int64)
    return h[:112].reshape(8, 14), h[112:112 + 392].reshape(8, 49), h[504:]


# ---- convenience wrappers -------------------------------------------------

def lanes_of(hand) -> "C.Array":
    """Hand (bridge_mc_b200.cards.Hand or 4 rank lists) -> c_uint16[4] rank masks."""
    suits = getattr(hand, "suits", hand)
    return (C.c_uint16 * 4)(*[sum(1 << r for r in s) for s in suits])


def assess(hand):
    """assess-hand: dict(lens, power, hcp, losers, balanced)."""
    out = (C.c_int * 11)()
    lib().orc_assess(lanes_of(hand), out)
    v = list(out)
    return {"lens": v[0:4], "power": v[4:8], "hcp": v[8], "losers": v[9], "balanced": bool(v[10])}


def choose_bid(scheme: int, hand) -> int:
    return lib().orc_choose_bid(scheme, lanes_of(hand))


def test_bid(scheme: int, k: int, hand) -> bool:
    return bool(lib().orc_test_bid(scheme, k, lanes_of(hand)))


def imps(diff: int):
    out = C.c_int()
    return out.value if lib().orc_imps(diff, C.byref(out)) == 0 else None


class Distgen:
    def __init__(self, cards=FULL_DECK, **spec):
        self._sp = make_spec(**spec)
        self._g = lib().orc_distgen_new(cards, C.byref(self._sp))

    def __del__(self):
        if getattr(self, "_g", None):
            lib().orc_distgen_free(self._g)
            self._g = None

    @property
    def npat(self):
        return lib().orc_distgen_npat(self._g)

    @property
    def total(self):
        return lib().orc_distgen_total(self._g)

    def items(self):
        p = C.c_uint()
        L = lib()
        for i in range(self.npat):
            w = L.orc_distgen_item(self._g, i, C.byref(p))
            yield w, p.value

    def pmf(self):
        hcp = (C.c_uint64 * 38)()
        lens = ((C.c_uint64 * 14) * 4)()
        lib().orc_distgen_pmf(self._g, hcp, lens)
        return list(hcp), [list(r) for r in lens]

    def rand_hand(self, state: "C.c_uint64"):
        return lib().orc_rand_hand(self._g, C.byref(state))

    def build_batch(self, n: int, seed: int = 1) -> np.ndarray:
        """n `build`s (one ranged seat at seat 0 + deal-all): uint64[n,4] hand masks (52-bit, id order)."""
        out = np.empty((n, 4), dtype=np.uint64)
        st = C.c_uint64(seed)
        lib().orc_build_batch(self._g, C.byref(st), n, out.ctypes.data)
        return out


def features_batch(seat_masks64: np.ndarray) -> np.ndarray:
    """uint64[m] 16-bit-lane hand masks -> int32[m, 11] assess-hand vectors."""
    m = np.ascontiguousarray(seat_masks64, dtype=np.uint64).reshape(-1)
    out = np.empty((m.size, 11), dtype=np.int32)
    lib().orc_features_batch(m.ctypes.data, m.size, out.ctypes.data)
    return out


def choose_bid_batch(scheme: int, seat_masks64: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(seat_masks64, dtype=np.uint64).reshape(-1)
    out = np.empty(m.size, dtype=np.int32)
    lib().orc_choose_bid_batch(scheme, m.ctypes.data, m.size, out.ctypes.data)
    return out


def gpu_deal_batch(seed: int, first: int, n: int):
    """CPU mirror of the GPU sampler: (records uint64[n,2], void uint8[n])."""
    rec = np.empty((n, 2), dtype=np.uint64)
    void = np.empty(n, dtype=np.uint8)
    lib().orc_gpu_deal_batch(seed, first, n, rec.ctypes.data, void.ctypes.data)
    return rec, void


def gpu_deal_fixed_batch(seed: int, first: int, n: int, fixed):
    """CPU mirror of the fixed-hand sampler; fixed = 4 lane-layout masks (0 = free seat)."""
    rec = np.empty((n, 2), dtype=np.uint64)
    void = np.empty(n, dtype=np.uint8)
    lib().orc_gpu_deal_fixed_batch(seed, first, n, (C.c_uint64 * 4)(*[int(f) for f in fixed]), rec.ctypes.data, void.ctypes.data)
    return rec, void


def gpu_deal_seatfirst_batch(seed: int, first: int, n: int, primary: int, hcp=(0, 37)):
    """CPU mirror of the seat-first sampler for a primary seat with HCP window `hcp`: (records, flags);
    flags bit0 / bit1 / bit2 = a Lemire check failed in stage A1 / A2 / B, bit 3 = the index passes stage A1."""
    rec = np.empty((n, 2), dtype=np.uint64)
    flags = np.empty(n, dtype=np.uint8)
    lib().orc_gpu_deal_seatfirst_batch(seed, first, n, primary, int(hcp[0]), int(hcp[1]), rec.ctypes.data, flags.ctypes.data)
    return rec, flags


def mask52_to_lanes64(m52: np.ndarray) -> np.ndarray:
    """52-bit id-order card masks -> 16-bit-lane 64-bit masks."""
    m = np.asarray(m52, dtype=np.uint64)
    out = np.zeros_like(m)
    for s in range(4):
        out |= ((m >> np.uint64(13 * s)) & np.uint64(0x1FFF)) << np.uint64(16 * s)
    return out


class ExactTable:
    """CPU mirror of the sampler's exact class table (mode 4) for a primary seat whose constraint is: inclusive ranges
    over the 11 features (dict: lens / power as 4-lists of (lo, hi), hcp, losers, balanced as (lo, hi)), optionally AND
    test-rule (kind 1) or choose-bid == k (kind 2) over one of the oracle's schemes."""

    def __init__(self, ranges=None, kind=0, scheme=0, k=0):
        lo, hi = [0] * 11, [13, 13, 13, 13, 10, 10, 10, 10, 37, 12, 1]
        ranges = ranges or {}
        for i, key in enumerate(("lens", "power")):
            for s, r in enumerate(ranges.get(key, [None] * 4)):
                if r is not None:
                    lo[4 * i + s], hi[4 * i + s] = r
        for i, key in ((8, "hcp"), (9, "losers"), (10, "balanced")):
            if ranges.get(key) is not None:
                lo[i], hi[i] = ranges[key]
        self._t = lib().orc_xtab_build(kind, scheme, k, (C.c_int * 11)(*lo), (C.c_int * 11)(*hi))

    def __del__(self):
        if getattr(self, "_t", None):
            lib().orc_xtab_free(self._t)
            self._t = None

    @property
    def weight(self):
        return lib().orc_xtab_weight(self._t)

    @property
    def classes(self):
        return lib().orc_xtab_classes(self._t)

    def deal_batch(self, seed: int, first: int, n: int, primary: int):
        """(records, flags): flags bit 0 = void in A1, bit 2 = void in B, bit 3 = the index passes stage A1"""
        rec = np.empty((n, 2), dtype=np.uint64)
        flags = np.empty(n, dtype=np.uint8)
        lib().orc_gpu_deal_exact_batch(self._t, seed, first, n, primary, rec.ctypes.data, flags.ctypes.data)
        return rec, flags
