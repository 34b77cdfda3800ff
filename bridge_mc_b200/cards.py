"""Cards, hands and the packed deal record (host-side mirror of the reference types).

Encoding follows the reference (bridge.cl:519-547): suit c=0 d=1 h=2 s=3, rank
0..12 = 2..A, card id = 13*suit + rank.  A `Hand` is what `build` returns at the
drop-in boundary (bridge.cl:1622-1633): four ascending rank lists (c d h s) and
an id.

Packed record (include/bridgemc.h, `bmc_record`): 128 bits = two 64-bit owner
bit-planes, one 16-bit lane per suit (bit r of lane s = card (s, r)).  Owner seat
of a card = 2*hi_plane_bit + lo_plane_bit, seats 0..3 = N E S W.
"""
from __future__ import annotations

import numpy as np

SUIT_CHARS = ("♣", "♦", "♥", "♠")
SUIT_SYMS = ("C", "D", "H", "S")
RANK_NAMES = ("2", "3", "4", "5", "6", "7", "8", "9", "10", "J", "Q", "K", "A")
SEATS = ("N", "E", "S", "W")
LANE_MASK = 0x1FFF1FFF1FFF1FFF


class Hand:
    """bridge.cl:1622-1633 class `hand`: `suits` = 4 ascending rank lists, `id`."""

    __slots__ = ("suits", "id")

    def __init__(self, suits, id=None):
        self.suits = [sorted(int(r) for r in s) for s in suits]
        self.id = id

    # --- conversions -----------------------------------------------------
    @classmethod
    def from_lanes(cls, lanes, id=None):
        """lanes: four 13-bit rank masks (c d h s)."""
        return cls([[r for r in range(13) if (int(m) >> r) & 1] for m in lanes], id)

    @classmethod
    def from_mask64(cls, mask, id=None):
        """mask: 64-bit, one 16-bit lane per suit (the record's lane layout)."""
        mask = int(mask)
        return cls.from_lanes([(mask >> (16 * s)) & 0x1FFF for s in range(4)], id)

    def lanes(self):
        return [sum(1 << r for r in s) for s in self.suits]

    def mask64(self) -> int:
        m = 0
        for s, lane in enumerate(self.lanes()):
            m |= lane << (16 * s)
        return m

    def cards(self):
        """card ids 13*suit+rank, ascending."""
        return [13 * s + r for s in range(4) for r in self.suits[s]]

    def __len__(self):
        return sum(len(s) for s in self.suits)

    def __eq__(self, other):
        return isinstance(other, Hand) and self.suits == other.suits

    def __hash__(self):
        return hash(tuple(tuple(s) for s in self.suits))

    def __repr__(self):          # bridge.cl:1699-1700 print-object -> HAND<...>
        return f"HAND<{print_hand(self.suits)}>"


def print_hand(suits) -> str:
    """bridge.cl:599-606 print-hand: "♣ ranks ♦ ranks ♥ ranks ♠ ranks", high to low."""
    return " ".join(f"{SUIT_CHARS[i]} " + "".join(RANK_NAMES[r] for r in sorted(s, reverse=True))
                    for i, s in enumerate(suits))


def str2hand(text: str) -> Hand:
    """bridge.cl:1645-1666 str2hand: "[Id: ]♣ KJ96 ♦ A8 ♥ AJ83 ♠ AK9"; suits in any
    order, `10` is two characters, a void is an empty field."""
    hand_id = None
    if ": " in text:
        hand_id, text = text.split(": ", 1)
    fields = text.split(" ")
    suits = [[], [], [], []]
    i = 0
    while i < len(fields):
        tok = fields[i]
        if tok not in SUIT_CHARS:
            i += 1
            continue
        sn = SUIT_CHARS.index(tok)
        nxt = fields[i + 1] if i + 1 < len(fields) else ""
        if nxt in SUIT_CHARS or nxt == "":
            suits[sn] = []
            i += 1 if nxt in SUIT_CHARS else 2
            continue
        ranks = []
        j = 0
        while j < len(nxt):
            if nxt[j] == "1":
                ranks.append(8)
                j += 2
            else:
                ranks.append("23456789_JQKA".index(nxt[j]))
                j += 1
        suits[sn] = ranks
        i += 2
    return Hand(suits, hand_id)


def str2deal(text: str):
    """bridge.cl:1675-1677 str2deal: one hand per line."""
    return [str2hand(line) for line in text.split("\n") if line.strip()]


def card(text: str):
    """bridge.cl:528-533 card: "S10" -> (3, 8)."""
    suit = SUIT_SYMS.index(text[0].upper())
    rank = RANK_NAMES.index(text[1:].upper())
    return (suit, rank)


# --- packed records ------------------------------------------------------

def pack_deal(hands) -> np.ndarray:
    """Four Hands (seat order N E S W) -> one record as uint64[2] = (lo_plane, hi_plane)."""
    lo = hi = 0
    seen = 0
    for seat, h in enumerate(hands):
        m = h.mask64() if isinstance(h, Hand) else int(h)
        if m & seen:
            raise ValueError("hands overlap")
        seen |= m
        if seat & 1:
            lo |= m
        if seat & 2:
            hi |= m
    if seen != LANE_MASK:
        raise ValueError("hands do not cover the deck")
    return np.array([lo, hi], dtype=np.uint64)


def pack_deals(deals) -> np.ndarray:
    """list of 4-hand deals -> uint64[n, 2] record array."""
    out = np.empty((len(deals), 2), dtype=np.uint64)
    for i, d in enumerate(deals):
        out[i] = pack_deal(d)
    return out


def seat_masks(records: np.ndarray) -> np.ndarray:
    """uint64[n,2] records -> uint64[n,4] per-seat 64-bit lane masks."""
    rec = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, 2)
    lo, hi = rec[:, 0], rec[:, 1]
    full = np.uint64(LANE_MASK)
    out = np.empty((rec.shape[0], 4), dtype=np.uint64)
    out[:, 0] = ~lo & ~hi & full
    out[:, 1] = lo & ~hi & full
    out[:, 2] = ~lo & hi & full
    out[:, 3] = lo & hi & full
    return out


def unpack_deal(record, ids=SEATS):
    """One record -> list of four Hands (N E S W)."""
    masks = seat_masks(np.asarray(record, dtype=np.uint64).reshape(1, 2))[0]
    return [Hand.from_mask64(int(m), ids[k] if ids else None) for k, m in enumerate(masks)]
