"""Thin object layer over the C ABI: Engine (bmc_ctx), Constraints, Scheme.

Everything computes on the GPU through libbridgemc.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import (BadRange, BridgeMcError, Contract, FEATURE_DTYPE, SeatSpec, Stats, Tallies, TALLY_ALL,
                   TALLY_WORDS, check)

DEFAULT_SEED = 0x6272696467656D63          # "bridgemc"

_KEYS = ("c", "d", "h", "s")


def _lo_hi(v):
    """number | (lo hi ...) -> (lo, hi) with None -> -1 (nil)."""
    if v is None:
        return -1, -1
    if isinstance(v, int):
        return v, v
    lo = -1 if v[0] is None else int(v[0])
    hi = -1 if len(v) < 2 or v[1] is None else int(v[1])
    return lo, hi


def seat_spec(hcp=None, c=None, d=None, h=None, s=None, losers=None, fixed=None) -> SeatSpec:
    """One seat's range def in the reference's keyword form
    (`:hcp (15 17) :s (5 nil (3 nil))`, README / rest.cl:250-256): each value is a
    number, (lo hi) or -- for suits -- (lo hi suit-hcp-range); None = nil."""
    sp = SeatSpec()
    sp.hcp[0], sp.hcp[1] = _lo_hi(hcp)
    for i, v in enumerate((c, d, h, s)):
        sp.len[i][0], sp.len[i][1] = _lo_hi(v)
        sh = v[2] if (v is not None and not isinstance(v, int) and len(v) > 2) else None
        sp.suit_hcp[i][0], sp.suit_hcp[i][1] = _lo_hi(sh)
    sp.losers[0], sp.losers[1] = _lo_hi(losers)
    if fixed is not None:
        sp.fixed = 1
        sp.fixed_mask = int(fixed.mask64() if hasattr(fixed, "mask64") else fixed)
    return sp


class Constraints:
    """bmc_constraints: per-seat ranges (seat_spec / dict / None) and rule forms (str / None)."""

    def __init__(self, specs=None, rules=None):
        L = _lib.load()
        arr = None
        if specs is not None:
            arr = (SeatSpec * 4)()
            for k in range(4):
                sp = specs[k] if k < len(specs) else None
                if sp is None:
                    sp = seat_spec()
                elif isinstance(sp, dict):
                    sp = seat_spec(**{kk.lower().lstrip(":"): vv for kk, vv in sp.items()})
                arr[k] = sp
        rarr = None
        if rules is not None:
            rarr = (C.c_char_p * 4)()
            for k in range(4):
                r = rules[k] if k < len(rules) else None
                rarr[k] = None if r is None else r.encode("utf-8")
        self._h = C.c_void_p()
        check(L.bmc_constraints_create(arr, rarr, C.byref(self._h)))

    MODE_AUTO, MODE_FULL, MODE_SEAT_FIRST, MODE_REFERENCE_ORDER, MODE_SEAT_FIRST_EXACT = 0, 1, 2, 3, 4

    def set_mode(self, mode: int):
        """bmc_constraints_set_mode: 0 auto, 1 full deal, 2 seat-first (see bridgemc.h)."""
        check(_lib.load().bmc_constraints_set_mode(self._h, mode))
        return self

    def set_reference_order(self, n_defs: int):
        """bmc_constraints_set_reference_order: sample the `n_defs` ranged seats one after the other, as `build` does."""
        check(_lib.load().bmc_constraints_set_reference_order(self._h, n_defs))
        return self

    def set_jit(self, on=True):
        """bmc_constraints_set_jit: True / 1 compile the rule forms into the kernel with NVRTC, False / 0 always interpret
        them, 2 (the library's default) switch by itself after a second of interpreted kernel time."""
        check(_lib.load().bmc_constraints_set_jit(self._h, int(on)))
        return self

    def set_scoring(self, contracts, vulnerable, pairs=()):
        """bmc_constraints_set_scoring: the contracts (Contract structs) and IMP pairs [(a, b), ...] that TALLY_SCORE
        tallies over the accepted deals."""
        nc = len(contracts)
        carr = (Contract * max(nc, 1))(*contracts)
        vul = np.ascontiguousarray(vulnerable, dtype=np.uint8)
        pr = np.ascontiguousarray(pairs, dtype=np.uint8).reshape(-1)
        check(_lib.load().bmc_constraints_set_scoring(self._h, carr, vul.ctypes.data if nc else None, nc,
                                                      pr.ctypes.data if pr.size else None, pr.size // 2))
        self.n_contracts, self.n_pairs = nc, pr.size // 2
        return self

    @property
    def jit_ms(self) -> float:
        return float(_lib.load().bmc_constraints_jit_ms(self._h))

    @property
    def mode(self) -> int:
        return _lib.load().bmc_constraints_mode(self._h)

    @property
    def primary(self) -> int:
        return _lib.load().bmc_constraints_primary(self._h)

    @property
    def primary_hcp(self):
        """(lo, hi): the HCP window of the primary seat that mode 2 orders the honour holdings by"""
        lo, hi = C.c_int32(), C.c_int32()
        check(_lib.load().bmc_constraints_primary_hcp(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    @property
    def primary_acceptance(self):
        """(hands, classes): the exact number of 13-card hands (of C(52,13)) the primary seat's constraint admits and the
        classes they fall into -- known after the first sampling call in mode 0 / 4 (device enumeration)."""
        hands, classes = C.c_uint64(), C.c_uint64()
        check(_lib.load().bmc_constraints_primary_acceptance(self._h, C.byref(hands), C.byref(classes)))
        return hands.value, classes.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().bmc_constraints_destroy(h)
            except Exception:          # interpreter shutdown: module globals may already be gone
                pass
            self._h = None


class Scheme:
    """bmc_scheme: a rule table `(((2 C) . FORM) ...)` for choose-bid."""

    def __init__(self, table_sexpr: str):
        self._h = C.c_void_p()
        check(_lib.load().bmc_scheme_create(table_sexpr.encode("utf-8"), C.byref(self._h)))
        self.size = _lib.load().bmc_scheme_size(self._h)

    def set_jit(self, on):
        """bmc_scheme_set_jit: 0 interpret the rows, 1 compile them into the filter kernel (NVRTC), 2 auto (default)"""
        check(_lib.load().bmc_scheme_set_jit(self._h, int(on)))
        return self

    @property
    def jit_ms(self) -> float:
        return float(_lib.load().bmc_scheme_jit_ms(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().bmc_scheme_destroy(h)
            except Exception:
                pass
            self._h = None


def tallies_to_dict(t: Tallies) -> dict:
    return {
        "raw": int(t.raw), "voided": int(t.voided), "accepted": int(t.accepted), "stored": int(t.stored),
        "hcp": np.ctypeslib.as_array(t.hcp).astype(np.int64)[:, :38].copy(),
        "len": np.ctypeslib.as_array(t.len).astype(np.int64).copy(),
        "shape": np.ctypeslib.as_array(t.shape).astype(np.int64)[:, :39].copy(),
        "tricks": np.ctypeslib.as_array(t.tricks).astype(np.int64).copy(),
        "imps": np.ctypeslib.as_array(t.imps).astype(np.int64).copy(),
        "imp_dropped": np.ctypeslib.as_array(t.imp_dropped).astype(np.int64).copy(),
    }


def tallies_from_words(words: np.ndarray) -> dict:
    """uint64[TALLY_WORDS] (e.g. an all-reduced device buffer copied back) -> dict."""
    w = np.ascontiguousarray(words, dtype=np.uint64)
    t = Tallies.from_buffer_copy(w.tobytes())
    return tallies_to_dict(t)


def sample_multi(devices, cons: Constraints | None = None, *, n_raw: int = 0, n_accept: int = 0, seed: int = DEFAULT_SEED,
                 first_index: int = 0, wave: int = 0, tally_mask: int = TALLY_ALL, cap: int = 0):
    """bmc_sample_multi: one call over several GPUs of the box (contiguous index shares, tallies summed in the library).
    Returns (records uint64[stored,2], tallies dict, Stats)."""
    L = _lib.load()
    devs = (C.c_int * len(devices))(*devices)
    records = np.empty((cap, 2), dtype=np.uint64) if cap else None
    t, st = Tallies(), Stats()
    check(L.bmc_sample_multi(devs, len(devices), cons._h if cons else None, seed, first_index, n_raw, n_accept, wave, tally_mask,
                             records.ctypes.data if cap else None, cap, C.byref(t), C.byref(st)))
    out = records[:st.stored] if cap else np.empty((0, 2), dtype=np.uint64)
    return out, tallies_to_dict(t), st


class Engine:
    """bmc_ctx on one device.  One Engine per calling thread (or share it: calls serialise)."""

    _default = None
    _default_lock = threading.Lock()

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        check(self._L.bmc_init(device, C.byref(self._h)))
        self.device = device

    @classmethod
    def default(cls) -> "Engine":
        with cls._default_lock:
            if cls._default is None:
                cls._default = cls(0)
            return cls._default

    def close(self):
        if getattr(self, "_h", None):
            self._L.bmc_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- sampling -----------------------------------------------------------
    def sample(self, cons: Constraints | None = None, *, n_raw: int = 0, n_accept: int = 0, seed: int = DEFAULT_SEED,
               first_index: int = 0, wave: int = 0, tally_mask: int = TALLY_ALL, cap: int | None = None,
               records: np.ndarray | None = None):
        """bmc_sample with host buffers.  Returns (records uint64[stored,2], tallies dict, Stats)."""
        if cap is None:
            cap = n_accept + (n_accept // 8) + 4096 if n_accept else n_raw
        if records is None and cap:
            records = np.empty((cap, 2), dtype=np.uint64)
        t, st = Tallies(), Stats()
        check(self._L.bmc_sample(self._h, cons._h if cons else None, seed, first_index, n_raw, n_accept, wave,
                                 tally_mask, records.ctypes.data if cap else None, cap, C.byref(t), C.byref(st)))
        out = records[:st.stored] if cap else np.empty((0, 2), dtype=np.uint64)
        return out, tallies_to_dict(t), st

    def sample_dev(self, cons: Constraints | None, records_ptr: int, cap: int, tallies_ptr: int, *, n_raw: int = 0,
                   n_accept: int = 0, seed: int = DEFAULT_SEED, first_index: int = 0, wave: int = 0,
                   tally_mask: int = TALLY_ALL) -> Stats:
        """bmc_sample_dev: device pointers (e.g. torch tensors' data_ptr())."""
        st = Stats()
        check(self._L.bmc_sample_dev(self._h, cons._h if cons else None, seed, first_index, n_raw, n_accept, wave,
                                     tally_mask, records_ptr or None, cap, tallies_ptr, C.byref(st)))
        return st

    def sample_indices(self, cons: Constraints | None = None, *, n_raw: int = 0, n_accept: int = 0, seed: int = DEFAULT_SEED,
                       first_index: int = 0, wave: int = 0, tally_mask: int = TALLY_ALL, cap: int | None = None,
                       offsets: np.ndarray | None = None):
        """bmc_sample_indices: like sample(), but every accepted deal comes back as the 32-bit offset of its index from
        first_index (uint32[stored]); expand() regenerates the records."""
        if cap is None:
            cap = n_accept + (n_accept // 8) + 4096 if n_accept else n_raw
        if offsets is None and cap:
            offsets = np.empty(cap, dtype=np.uint32)
        t, st = Tallies(), Stats()
        check(self._L.bmc_sample_indices(self._h, cons._h if cons else None, seed, first_index, n_raw, n_accept, wave,
                                         tally_mask, offsets.ctypes.data if cap else None, cap, C.byref(t), C.byref(st)))
        out = offsets[:st.stored] if cap else np.empty(0, dtype=np.uint32)
        return out, tallies_to_dict(t), st

    def sample_packed(self, cons: Constraints | None = None, *, n_raw: int = 0, n_accept: int = 0, seed: int = DEFAULT_SEED,
                      first_index: int = 0, wave: int = 0, tally_mask: int = TALLY_ALL, cap_bytes: int | None = None,
                      out: np.ndarray | None = None):
        """bmc_sample_packed: like sample_indices(), but the accepted indices come back as gap bytes (about one byte per
        deal; include/bridgemc.h).  Returns (uint8[n_bytes], tallies, stats); unpack_indices() gives the offsets."""
        if out is not None:
            cap_bytes = out.nbytes
        elif cap_bytes is None:
            expect = n_accept + n_accept // 4 if n_accept else n_raw
            cap_bytes = int(expect * 1.2) + (max(n_raw, 1 << 26) >> 16) * 8 + (1 << 20)
        if out is None:
            out = np.empty(cap_bytes, dtype=np.uint8)
        t, st, nb = Tallies(), Stats(), C.c_uint64()
        check(self._L.bmc_sample_packed(self._h, cons._h if cons else None, seed, first_index, n_raw, n_accept, wave, tally_mask,
                                        out.ctypes.data, cap_bytes, C.byref(nb), C.byref(t), C.byref(st)))
        return out[:nb.value], tallies_to_dict(t), st

    def unpack_indices(self, packed: np.ndarray, cap: int | None = None) -> np.ndarray:
        """bmc_unpack_indices: gap bytes -> uint32 offsets from first_index (chunk order; sort for index order)."""
        b = np.ascontiguousarray(packed, dtype=np.uint8)
        n = C.c_uint64()
        if cap is None:
            check(self._L.bmc_unpack_indices(b.ctypes.data, b.size, None, 0, C.byref(n)))
            cap = n.value
        offs = np.empty(cap, dtype=np.uint32)
        check(self._L.bmc_unpack_indices(b.ctypes.data, b.size, offs.ctypes.data, cap, C.byref(n)))
        return offs[:n.value]

    def sample_indices_dev(self, cons: Constraints | None, offsets_ptr: int, cap: int, tallies_ptr: int, *, n_raw: int = 0,
                           n_accept: int = 0, seed: int = DEFAULT_SEED, first_index: int = 0, wave: int = 0,
                           tally_mask: int = TALLY_ALL) -> Stats:
        st = Stats()
        check(self._L.bmc_sample_indices_dev(self._h, cons._h if cons else None, seed, first_index, n_raw, n_accept, wave,
                                             tally_mask, offsets_ptr or None, cap, tallies_ptr, C.byref(st)))
        return st

    def expand(self, cons: Constraints | None, offsets: np.ndarray, *, seed: int = DEFAULT_SEED, first_index: int = 0) -> np.ndarray:
        """bmc_expand: index offsets -> records uint64[n,2]."""
        off = np.ascontiguousarray(offsets, dtype=np.uint32).reshape(-1)
        out = np.empty((off.size, 2), dtype=np.uint64)
        check(self._L.bmc_expand(self._h, cons._h if cons else None, seed, first_index, off.ctypes.data if off.size else None,
                                 off.size, out.ctypes.data if off.size else None))
        return out

    def expand_dev(self, cons: Constraints | None, offsets_ptr: int, n: int, records_ptr: int, *, seed: int = DEFAULT_SEED,
                   first_index: int = 0):
        check(self._L.bmc_expand_dev(self._h, cons._h if cons else None, seed, first_index, offsets_ptr, n, records_ptr))

    def set_stream(self, cuda_stream: int):
        check(self._L.bmc_set_stream(self._h, cuda_stream or None))

    # ---- filter ---------------------------------------------------------------
    def filter(self, records: np.ndarray, cons: Constraints | None = None, scheme: Scheme | None = None,
               want_features: bool = True):
        """bmc_filter: (accept uint8[n], features structured[n,4])."""
        rec = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, 2)
        n = rec.shape[0]
        acc = np.empty(n, dtype=np.uint8)
        feats = np.empty((n, 4), dtype=FEATURE_DTYPE) if want_features else None
        check(self._L.bmc_filter(self._h, cons._h if cons else None, scheme._h if scheme else None,
                                 rec.ctypes.data if n else None, n, acc.ctypes.data,
                                 feats.ctypes.data if want_features else None))
        return acc, feats

    def filter_dev(self, records_ptr: int, n: int, cons: Constraints | None = None, scheme: Scheme | None = None,
                   accept_ptr: int = 0, feats_ptr: int = 0):
        """bmc_filter_dev: device pointers, asynchronous on the context's stream"""
        check(self._L.bmc_filter_dev(self._h, cons._h if cons else None, scheme._h if scheme else None, records_ptr, n,
                                     accept_ptr or None, feats_ptr or None))

    def auction(self, records: np.ndarray, openings: Scheme, responses, dealer: int = 0) -> np.ndarray:
        """bmc_auction: class `bidding` / `next` (bidding.cl:457-480) over a deal set.  `responses` = one Scheme (or None)
        per row of `openings`.  Returns a structured array (opener, opening, response, status)."""
        rec = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, 2)
        n = rec.shape[0]
        arr = (C.c_void_p * max(1, len(responses)))(*[(r._h if r is not None else None) for r in responses])
        out = np.empty(n, dtype=_lib.AUCTION_DTYPE)
        check(self._L.bmc_auction(self._h, openings._h, arr, dealer, rec.ctypes.data if n else None, n, out.ctypes.data))
        return out

    # ---- scoring ----------------------------------------------------------------
    def score(self, contracts, vulnerable, tricks: np.ndarray) -> np.ndarray:
        """bmc_score: tricks uint8[n, n_contracts] -> base scores int32[n, n_contracts]."""
        nc = len(contracts)
        carr = (Contract * nc)(*contracts)
        vul = np.ascontiguousarray(vulnerable, dtype=np.uint8)
        tr = np.ascontiguousarray(tricks, dtype=np.uint8).reshape(-1, nc)
        out = np.empty(tr.shape, dtype=np.int32)
        check(self._L.bmc_score(self._h, carr, vul.ctypes.data, nc, tr.ctypes.data, tr.shape[0], out.ctypes.data))
        return out

    def match_points(self, a: np.ndarray, b: np.ndarray):
        """bmc_match_points: (imps int8[n], status uint8[n])."""
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1)
        imps = np.empty(a.size, dtype=np.int8)
        status = np.empty(a.size, dtype=np.uint8)
        check(self._L.bmc_match_points(self._h, a.ctypes.data, b.ctypes.data, a.size, imps.ctypes.data, status.ctypes.data))
        return imps, status

    def score_tally(self, contracts, vulnerable, pairs, n: int, seed: int = DEFAULT_SEED, first_index: int = 0):
        nc = len(contracts)
        carr = (Contract * nc)(*contracts)
        vul = np.ascontiguousarray(vulnerable, dtype=np.uint8)
        pr = np.ascontiguousarray(pairs, dtype=np.uint8).reshape(-1)
        npairs = pr.size // 2
        th = np.zeros((nc, 14), dtype=np.uint64)
        ih = np.zeros((max(npairs, 1), 49), dtype=np.uint64)
        dropped = np.zeros(max(npairs, 1), dtype=np.uint64)
        tab = np.zeros((nc, 14), dtype=np.int32)
        check(self._L.bmc_score_tally(self._h, carr, vul.ctypes.data, nc, pr.ctypes.data if npairs else None, npairs,
                                      seed, first_index, n, th.ctypes.data, ih.ctypes.data, dropped.ctypes.data, tab.ctypes.data))
        self.last_imp_dropped = dropped[:npairs]
        return th, ih[:npairs], tab

    # ---- play-deal: the trick estimator -------------------------------------------------
    def play_deal(self, records: np.ndarray, trump, declarer, arena_kb: int = 0) -> np.ndarray:
        """bmc_play_deal: the reference's `play-deal` (bridge.cl:2961) on every record, one GPU thread per deal.
        `trump` (-1 = no trump, 0..3 = c d h s) and `declarer` (seat 0..3) are scalars or per-deal arrays.
        Returns a structured array (_lib.PLAY_DTYPE): status, n_tricks, score (defence, declarer), cards, winner."""
        rec = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, 2)
        n = rec.shape[0]
        per = int(np.ndim(trump) > 0 or np.ndim(declarer) > 0)
        tr = np.ascontiguousarray(np.broadcast_to(np.asarray(trump, dtype=np.int8), (n,) if per else (1,)))
        de = np.ascontiguousarray(np.broadcast_to(np.asarray(declarer, dtype=np.uint8), (n,) if per else (1,)))
        out = np.empty(n, dtype=_lib.PLAY_DTYPE)
        if n:
            check(self._L.bmc_play_deal(self._h, rec.ctypes.data, n, tr.ctypes.data, de.ctypes.data, per, arena_kb, out.ctypes.data))
        return out

    def play_hands(self, hands: np.ndarray, trump, arena_kb: int = 0) -> np.ndarray:
        """bmc_play_hands: `hands` uint64[n, 4] lane masks (declarer, left, dummy, right), possibly partial deals."""
        hm = np.ascontiguousarray(hands, dtype=np.uint64).reshape(-1, 4)
        n = hm.shape[0]
        per = int(np.ndim(trump) > 0)
        tr = np.ascontiguousarray(np.broadcast_to(np.asarray(trump, dtype=np.int8), (n,) if per else (1,)))
        out = np.empty(n, dtype=_lib.PLAY_DTYPE)
        if n:
            check(self._L.bmc_play_hands(self._h, hm.ctypes.data, n, tr.ctypes.data, per, arena_kb, out.ctypes.data))
        return out

    def play_deal_dev(self, records_ptr: int, n: int, trump_ptr: int, declarer_ptr: int, per_deal: bool, out_ptr: int, arena_kb: int = 0):
        """bmc_play_deal_dev: device pointers in, device pointer out (72-byte results)."""
        check(self._L.bmc_play_deal_dev(self._h, records_ptr, n, trump_ptr, declarer_ptr, int(per_deal), arena_kb, out_ptr))

    def best_tricks(self, records: np.ndarray, trump, declarer, arena_kb: int = 0) -> np.ndarray:
        """bmc_best_tricks: `best-tricks` (bridge.cl:3027-3028) per record; 255 where play-deal gave no outcome."""
        rec = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, 2)
        n = rec.shape[0]
        per = int(np.ndim(trump) > 0 or np.ndim(declarer) > 0)
        tr = np.ascontiguousarray(np.broadcast_to(np.asarray(trump, dtype=np.int8), (n,) if per else (1,)))
        de = np.ascontiguousarray(np.broadcast_to(np.asarray(declarer, dtype=np.uint8), (n,) if per else (1,)))
        out = np.empty(n, dtype=np.uint8)
        if n:
            check(self._L.bmc_best_tricks(self._h, rec.ctypes.data, n, tr.ctypes.data, de.ctypes.data, per, arena_kb, out.ctypes.data))
        return out

    # ---- hist-base on the device -----------------------------------------------------
    def hist_base(self, records: np.ndarray, measures, cap: int = 1 << 20):
        """bmc_hist_base: `measures` = [(kind, seat, suit), ...] (kinds: _lib.M_*).  Returns hist-base rows
        [[[m1, .., mk], count], ...] in first-seen order of `records` (bridge.cl:1325-1330)."""
        rec = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, 2)
        k = len(measures)
        marr = (_lib.Measure * k)(*[_lib.Measure(int(m[0]), int(m[1]), int(m[2]) if len(m) > 2 else 0, 0) for m in measures])
        tuples = np.empty((cap, k), dtype=np.uint8)
        counts = np.empty(cap, dtype=np.uint64)
        nu = C.c_uint64()
        check(self._L.bmc_hist_base(self._h, rec.ctypes.data if rec.shape[0] else None, rec.shape[0], marr, k,
                                    tuples.ctypes.data, counts.ctypes.data, cap, C.byref(nu)))
        return [[[int(x) for x in tuples[i]], int(counts[i])] for i in range(nu.value)]

    # ---- distgen on the device --------------------------------------------------------
    def distgen(self, spec=None, cards: int = _lib.LANE_MASK, cap: int = 65536):
        """bmc_distgen: the reference's exact enumerator (bridge.cl:921-1045) for one range spec (seat_spec / dict / None)
        over the deck `cards` (lane mask).  Returns dict(patterns uint16[n], weights uint64[n], total, hcp_pmf[38],
        len_pmf[4,14]); pattern bit 15 = club J ... bit 0 = spade A."""
        if isinstance(spec, dict):
            spec = seat_spec(**{kk.lower().lstrip(":"): vv for kk, vv in spec.items()})
        pats = np.zeros(cap, dtype=np.uint16)
        wts = np.zeros(cap, dtype=np.uint64)
        hcp = np.zeros(38, dtype=np.uint64)
        lens = np.zeros((4, 14), dtype=np.uint64)
        n, total = C.c_uint32(), C.c_uint64()
        check(self._L.bmc_distgen(self._h, cards, C.byref(spec) if spec is not None else None, pats.ctypes.data, wts.ctypes.data,
                                  cap, C.byref(n), C.byref(total), hcp.ctypes.data, lens.ctypes.data))
        return {"patterns": pats[:n.value], "weights": wts[:n.value], "total": total.value, "hcp_pmf": hcp, "len_pmf": lens}

    # ---- measurement --------------------------------------------------------------
    def int_peak(self):
        ops = C.c_double()
        ms = C.c_float()
        check(self._L.bmc_int_peak(self._h, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value

    def launch_count(self) -> int:
        return int(self._L.bmc_launch_count(self._h))
