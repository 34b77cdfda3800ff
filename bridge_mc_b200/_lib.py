"""ctypes binding of libbridgemc.so (include/bridgemc.h) -- the same C ABI the
Lisp CFFI shim binds.  There is no fallback: if the shared
library is missing or no CUDA device is present, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbridgemc.so")

BMC_OK, BMC_E_INVALID, BMC_E_CUDA, BMC_E_BAD_RANGE, BMC_E_PARSE, BMC_E_NOMEM, BMC_E_CAPACITY = 0, -1, -2, -3, -4, -5, -6
TALLY_HCP, TALLY_LEN, TALLY_SHAPE, TALLY_ALL, TALLY_SCORE = 1, 2, 4, 7, 8
HCP_BINS, LEN_BINS, SHAPE_BINS = 40, 14, 40
SCORE_CONTRACTS, SCORE_PAIRS = 8, 8
LANE_MASK = 0x1FFF1FFF1FFF1FFF


class BridgeMcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbridgemc error {code}: {msg}")
        self.code = code


class BadRange(BridgeMcError):
    """The reference's `bad-range` condition (bridge.cl:1109-1123)."""


class SeatSpec(C.Structure):
    _fields_ = [("hcp", C.c_int8 * 2), ("len", (C.c_int8 * 2) * 4), ("suit_hcp", (C.c_int8 * 2) * 4),
                ("losers", C.c_int8 * 2), ("fixed", C.c_uint8), ("reserved", C.c_uint8 * 3),
                ("fixed_mask", C.c_uint64)]


class Tallies(C.Structure):
    _fields_ = [("raw", C.c_uint64), ("voided", C.c_uint64), ("accepted", C.c_uint64), ("stored", C.c_uint64),
                ("hcp", (C.c_uint64 * HCP_BINS) * 4), ("len", ((C.c_uint64 * LEN_BINS) * 4) * 4),
                ("shape", (C.c_uint64 * SHAPE_BINS) * 4),
                ("tricks", (C.c_uint64 * 14) * SCORE_CONTRACTS), ("imps", (C.c_uint64 * 49) * SCORE_PAIRS),
                ("imp_dropped", C.c_uint64 * SCORE_PAIRS)]


TALLY_WORDS = C.sizeof(Tallies) // 8


class Stats(C.Structure):
    _fields_ = [("raw", C.c_uint64), ("voided", C.c_uint64), ("accepted", C.c_uint64), ("stored", C.c_uint64),
                ("waves", C.c_uint32), ("launches", C.c_uint32), ("kernel_ms", C.c_float), ("total_ms", C.c_float)]


class Measure(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("seat", C.c_uint8), ("suit", C.c_uint8), ("reserved", C.c_uint8)]


M_HCP, M_LEN, M_SUIT_HCP, M_LOSERS, M_BALANCED, M_LINE_HCP, M_TRUMP_LEN, M_SHAPE, M_LONGEST = range(9)


class Contract(C.Structure):
    _fields_ = [("level", C.c_int8), ("suit", C.c_int8), ("declarer", C.c_int8), ("dbl", C.c_int8)]


FEATURE_DTYPE = np.dtype([("len", np.uint8, (4,)), ("suit_hcp", np.uint8, (4,)), ("hcp", np.uint8),
                          ("losers", np.uint8), ("balanced", np.uint8), ("bid", np.int8)])
assert FEATURE_DTYPE.itemsize == 12
PLAY_DTYPE = np.dtype([("status", np.uint8), ("n_tricks", np.uint8), ("score", np.uint8, (2,)), ("cards", np.uint8, (13, 4)),
                       ("winner", np.uint8, (13,)), ("reserved", np.uint8, (3,))])
assert PLAY_DTYPE.itemsize == 72
AUCTION_DTYPE = np.dtype([("opener", np.int8), ("opening", np.int8), ("response", np.int8), ("status", np.int8)])

EXPORTS = [
    "bmc_init", "bmc_shutdown", "bmc_last_error", "bmc_version", "bmc_host_alloc", "bmc_host_free",
    "bmc_shape_table", "bmc_constraints_create", "bmc_constraints_destroy", "bmc_constraints_set_mode", "bmc_constraints_set_reference_order", "bmc_constraints_set_jit", "bmc_constraints_jit_ms", "bmc_constraints_mode", "bmc_constraints_primary", "bmc_constraints_primary_hcp",
    "bmc_auction", "bmc_auction_dev", "bmc_scheme_create", "bmc_scheme_set_jit", "bmc_scheme_jit_ms", "bmc_filter_dev",
    "bmc_constraints_set_scoring", "bmc_constraints_primary_acceptance", "bmc_distgen", "bmc_sample_indices", "bmc_sample_indices_dev", "bmc_expand", "bmc_expand_dev", "bmc_sample_multi",
    "bmc_scheme_size", "bmc_scheme_destroy", "bmc_sample", "bmc_sample_dev", "bmc_filter", "bmc_score",
    "bmc_match_points", "bmc_score_tally", "bmc_hist_base", "bmc_int_peak", "bmc_launch_count", "bmc_set_stream",
    "bmc_play_deal", "bmc_play_deal_dev", "bmc_play_hands", "bmc_best_tricks", "bmc_sample_packed", "bmc_unpack_indices",
]

_lib = None


def load():
    """dlopen libbridgemc.so and declare the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(bridge_mc_b200/csrc/build.sh).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    L.bmc_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.bmc_shutdown.argtypes = [vp]
    L.bmc_shutdown.restype = None
    L.bmc_last_error.restype = C.c_char_p
    L.bmc_version.restype = C.c_char_p
    L.bmc_host_alloc.argtypes = [C.c_size_t]
    L.bmc_host_alloc.restype = vp
    L.bmc_host_free.argtypes = [vp]
    L.bmc_host_free.restype = None
    L.bmc_shape_table.argtypes = [vp]
    L.bmc_shape_table.restype = None
    L.bmc_constraints_create.argtypes = [C.POINTER(SeatSpec), C.POINTER(C.c_char_p), C.POINTER(vp)]
    L.bmc_constraints_destroy.argtypes = [vp]
    L.bmc_constraints_destroy.restype = None
    L.bmc_constraints_set_mode.argtypes = [vp, C.c_int]
    L.bmc_constraints_set_reference_order.argtypes = [vp, C.c_int]
    L.bmc_constraints_set_jit.argtypes = [vp, C.c_int]
    L.bmc_constraints_jit_ms.argtypes = [vp]
    L.bmc_constraints_jit_ms.restype = C.c_float
    L.bmc_constraints_mode.argtypes = [vp]
    L.bmc_constraints_primary.argtypes = [vp]
    L.bmc_constraints_primary_hcp.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.bmc_scheme_create.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.bmc_scheme_size.argtypes = [vp]
    L.bmc_scheme_destroy.argtypes = [vp]
    L.bmc_scheme_destroy.restype = None
    L.bmc_sample.argtypes = [vp, vp, u64, u64, u64, u64, u64, u32, vp, u64, C.POINTER(Tallies), C.POINTER(Stats)]
    L.bmc_sample_dev.argtypes = [vp, vp, u64, u64, u64, u64, u64, u32, vp, u64, vp, C.POINTER(Stats)]
    L.bmc_constraints_primary_acceptance.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.bmc_distgen.argtypes = [vp, u64, C.POINTER(SeatSpec), vp, vp, u32, C.POINTER(u32), C.POINTER(u64), vp, vp]
    L.bmc_constraints_set_scoring.argtypes = [vp, C.POINTER(Contract), vp, u32, vp, u32]
    L.bmc_sample_indices.argtypes = [vp, vp, u64, u64, u64, u64, u64, u32, vp, u64, C.POINTER(Tallies), C.POINTER(Stats)]
    L.bmc_sample_indices_dev.argtypes = [vp, vp, u64, u64, u64, u64, u64, u32, vp, u64, vp, C.POINTER(Stats)]
    L.bmc_expand.argtypes = [vp, vp, u64, u64, vp, u64, vp]
    L.bmc_expand_dev.argtypes = [vp, vp, u64, u64, vp, u64, vp]
    L.bmc_sample_multi.argtypes = [C.POINTER(C.c_int), C.c_int, vp, u64, u64, u64, u64, u64, u32, vp, u64, C.POINTER(Tallies),
                                   C.POINTER(Stats)]
    L.bmc_filter.argtypes = [vp, vp, vp, vp, u64, vp, vp]
    L.bmc_filter_dev.argtypes = [vp, vp, vp, vp, u64, vp, vp]
    L.bmc_auction.argtypes = [vp, vp, C.POINTER(vp), C.c_int, vp, u64, vp]
    L.bmc_auction_dev.argtypes = [vp, vp, C.POINTER(vp), C.c_int, vp, u64, vp]
    L.bmc_scheme_set_jit.argtypes = [vp, C.c_int]
    L.bmc_scheme_jit_ms.argtypes = [vp]
    L.bmc_scheme_jit_ms.restype = C.c_float
    L.bmc_score.argtypes = [vp, C.POINTER(Contract), vp, u32, vp, u64, vp]
    L.bmc_match_points.argtypes = [vp, vp, vp, u64, vp, vp]
    L.bmc_score_tally.argtypes = [vp, C.POINTER(Contract), vp, u32, vp, u32, u64, u64, u64, vp, vp, vp, vp]
    L.bmc_hist_base.argtypes = [vp, vp, u64, C.POINTER(Measure), u32, vp, vp, u64, C.POINTER(u64)]
    L.bmc_int_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.bmc_launch_count.argtypes = [vp]
    L.bmc_launch_count.restype = u64
    L.bmc_set_stream.argtypes = [vp, vp]
    L.bmc_play_deal.argtypes = [vp, vp, u64, vp, vp, C.c_int, u32, vp]
    L.bmc_play_deal_dev.argtypes = [vp, vp, u64, vp, vp, C.c_int, u32, vp]
    L.bmc_best_tricks.argtypes = [vp, vp, u64, vp, vp, C.c_int, u32, vp]
    L.bmc_play_hands.argtypes = [vp, vp, u64, vp, C.c_int, u32, vp]
    L.bmc_sample_packed.argtypes = [vp, vp, u64, u64, u64, u64, u64, u32, vp, u64, C.POINTER(u64), C.POINTER(Tallies), C.POINTER(Stats)]
    L.bmc_unpack_indices.argtypes = [vp, u64, vp, u64, C.POINTER(u64)]
    _lib = L
    return L


def check(rc: int):
    if rc == BMC_OK:
        return
    msg = load().bmc_last_error().decode("utf-8", "replace")
    if rc == BMC_E_BAD_RANGE:
        raise BadRange(rc, msg)
    raise BridgeMcError(rc, msg)


def shape_table() -> np.ndarray:
    out = np.zeros((39, 4), dtype=np.uint8)
    load().bmc_shape_table(out.ctypes.data)
    return out
