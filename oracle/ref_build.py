"""Reference-algorithm `build` with any number of ranged seats (bridge.cl:1292-1319), for distribution checks of
the GPU's reference-order mode.  TEST INFRASTRUCTURE ONLY, and self-contained: the distgen enumeration and sampling,
the per-seat narrowing (infer-hcp-range / infer-suit-ranges / infer-distgen-params, bridge.cl:1126-1269) and the
build loop are all the C restatement in bmc_oracle.c -- nothing here imports the product package.  The C infer-* is
pinned by the reference's own 14 KATs (tests/test_oracle_kats.py)."""
from __future__ import annotations

import numpy as np

from . import oracle as O


def ref_build_batch(consts, defs, n: int, seed: int = 1, threads: int = 0):
    """n builds.  consts: 52-bit id-order masks; defs: list of dicts ({'hcp': [lo, hi], 'c': [lo, hi, [slo, shi]], ..}).
    Returns uint64[m,4] hands in `build`'s order (consts, sampled seats, deal-all seats) for the m builds that did not
    signal an error, and the number that did."""
    hands, cnt = O.ref_build_batch(consts, defs, n, seed, threads)
    good = cnt > 0
    return hands[good], int((~good).sum())
