// playdeal.cuh -- the reference's trick estimator `play-deal` (bridge.cl:1497-2963), one deal per warp.
//
// `play-deal trump declarer left dummy right` is a heuristic search: at every lead it tries three ways of playing on
// (`play-strategy`, `mk-route`, a plain trick in the leader's first suit, bridge.cl:2949-2959), each of which rests on
// suit-by-suit look-aheads (`suit-tricks` 2279, `play-routes` 2476, `immediate-tricks` 2705) built from one primitive,
// `simple-trick` (2269): second hand covers or finesses, third hand finesses, fourth hand beats or plays low.  The answer
// depends on evaluation order and on state the reference mutates in place and shares between branches of the search
// (routes are popped by `reproduce`, 2372; the ranks to hunt are popped by `play-strategy`, 2863; a play-com's guards
// are overwritten, 2614), so this file keeps the reference's objects as records in a per-deal arena, referenced by
// offset, created exactly where the reference calls make-instance, and keeps its evaluation order.
//
// Layout
//   card    one byte, suit << 4 | rank (suit c d h s = 0..3, rank 0..12 = 2..A); 0xFF = nil
//   hand    four 13-bit rank masks + the seat it descends from (the reference compares hand ids with `eq`)
//   trick   four cards in play order, packed in one 32-bit word inside routes
//   arena   per-deal bump allocator (global memory); the search releases it when it backtracks
// Every helper that picks a card is bit arithmetic on a 13-bit mask (lowest = ffs, highest = 31 - clz, first above r =
// ffs(mask >> (r + 1))).  No recursion: the four recursive searches of the reference run on explicit stacks.
//
// Compiled by nvcc into libbridgemc.so (kernel `play_deal_kernel`, bridgemc.cu) and -- the same text -- by g++ into
// tests/playdeal_host_selftest, a test-only binary with which the CPU suite checks this file without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PD_HD __host__ __device__ __forceinline__
#define PD_HDN __host__ __device__ __noinline__
#define PD_ROLLED _Pragma("unroll 1")      /* the estimator is bound by instruction fetch: loops stay rolled */
#else
#define PD_HD inline
#define PD_HDN
#define PD_ROLLED
#endif

// On the device the 32 lanes of a warp run one deal in lock step (play_deal_kernel, bridgemc.cu): the scalar logic is
// executed redundantly by every lane on identical data, and the bulk list operations below (copy, scan, compaction of
// trick lists) are split across the lanes.  On the host there is one lane.
#if defined(__CUDA_ARCH__)
#define PD_LANE ((int)(threadIdx.x & 31u))
#define PD_LANES 32
#define PD_SYNC() __syncwarp()
#else
#define PD_LANE 0
#define PD_LANES 1
#define PD_SYNC() ((void)0)
#endif

// -DPD_PROFILE (diagnostic builds only): SM cycles per phase of a deal, summed over the deals of a launch
#if defined(PD_PROFILE) && defined(__CUDACC__)
__device__ unsigned long long g_pd_prof[8];
#endif
#if defined(PD_PROFILE) && defined(__CUDA_ARCH__)
#define PD_PROF_T0() long long pd_prof_t0 = clock64()
#define PD_PROF_ADD(slot) do { const long long pd_now = clock64(); if ((threadIdx.x & 31u) == 0u) atomicAdd(&g_pd_prof[slot], (unsigned long long)(pd_now - pd_prof_t0)); pd_prof_t0 = pd_now; } while (0)
#else
#define PD_PROF_T0() ((void)0)
#define PD_PROF_ADD(slot) ((void)0)
#endif

namespace pd {

typedef uint8_t Card;
constexpr Card NOCARD = 0xFF;
constexpr int MAXR = 40;         // routes per list
constexpr int MAXG = 8;          // guards per play-com
constexpr int MAXT = 6;          // targets per strategy
constexpr int MAXD = 14;         // search depth (13 tricks + the root)

enum Status : int { PD_OK = 0, PD_E_ARENA = 1, PD_E_LISP = 2, PD_E_CAP = 3, PD_E_NIL = 4 };

PD_HD Card mk(int suit, int rank) { return (Card)((suit << 4) | rank); }
PD_HD int csuit(Card c) { return c >> 4; }
PD_HD int crank(Card c) { return c & 15; }
PD_HD int cidx(Card c) { return c == NOCARD ? 52 : csuit(c) * 13 + crank(c); }
PD_HD uint32_t pack4(const Card *c) { return (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24); }
PD_HD Card tcard(uint32_t t, int i) { return (Card)(t >> (8 * i)); }
PD_HD int lowbit(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}
PD_HD int highbit(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)m);
#else
    return 31 - __builtin_clz(m);
#endif
}
PD_HD int popc(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}
// ranks of `m` strictly above r (r may be -1)
PD_HD uint32_t above(uint32_t m, int r) { return r < 0 ? m : (r >= 12 ? 0u : (m & ~((2u << r) - 1u))); }

// A hand is ONE 64-bit word: suit s in bits [16 s, 16 s + 13), the seat it descends from in bits 62..63 -- copying a hand
// is a register move and taking a card one AND.  `h.s[suit]` reads a suit's 13-bit rank mask.
struct Suits {
    uint64_t m;
    PD_HD uint32_t operator[](int suit) const { return (uint32_t)(m >> (16 * suit)) & 0x1FFFu; }
};
struct Hand {
    Suits s;
    PD_HD int id() const { return (int)(s.m >> 62) & 3; }
    PD_HD void set_id(int k) { s.m = (s.m & ~(3ull << 62)) | ((uint64_t)(k & 3) << 62); }
    PD_HD void set(int suit, uint32_t mask) { s.m = (s.m & ~(0x1FFFull << (16 * suit))) | ((uint64_t)(mask & 0x1FFFu) << (16 * suit)); }
};
PD_HD bool hand_empty(const Hand &h) { return (h.s.m & 0x1FFF1FFF1FFF1FFFull) == 0; }

struct alignas(16) Outcome {     // bridge.cl:2190-2196; winners = card wno[i] of tricks[i]
    uint8_t n;                   // tricks played
    uint8_t roll;
    uint8_t score[2];
    uint8_t valid;               // 0 = nil (no outcome)
    uint8_t wno[13];
    uint32_t tricks[13];         // four cards in play order, one byte each (pack4 / tcard)
    Hand rem[4];
};

struct Ctx {                     // per-deal state
    uint8_t *arena;
    uint32_t cap, top;
    int status;
    int trump;                   // -1 = no trump
    uint8_t *sstk;               // stack of the functions' temporaries: shared memory on the device (one per warp), so
    uint32_t scap, stop, speak;  // that nothing indexed dynamically lives in (lane-interleaved) local memory
    uint32_t pos;                // bucket table of the deal's position cache (grows down from the arena's end); 0 = none
    uint32_t empties;            // four play-coms without routes or guards, one per hand on lead; 0 = not made yet
};

PD_HD uint32_t arena_alloc(Ctx &cx, uint32_t bytes)
{
    bytes = (bytes + 3u) & ~3u;
    if (cx.top + bytes > cx.cap) {
        cx.status = cx.status ? cx.status : PD_E_ARENA;
        return 0;                // offset 0 is a scratch area that is always valid
    }
    const uint32_t at = cx.top;
    cx.top += bytes;
    return at;
}
template <class T> PD_HD T *at(Ctx &cx, uint32_t off) { return reinterpret_cast<T *>(cx.arena + off); }

PD_HD void fail(Ctx &cx, int code) { if (!cx.status) cx.status = code; }

// temporaries: `T &x = *tmp<T>(cx)` inside a function that opens with `TmpFrame tf(cx);`
template <class T> PD_HD T *tmp(Ctx &cx, uint32_t count = 1)
{
    const uint32_t bytes = ((uint32_t)sizeof(T) * count + 15u) & ~15u;
    if (cx.stop + bytes > cx.scap) { fail(cx, PD_E_CAP); return reinterpret_cast<T *>(cx.sstk); }   // scap >= any single request
    T *p = reinterpret_cast<T *>(cx.sstk + cx.stop);
    cx.stop += bytes;
    if (cx.stop > cx.speak) cx.speak = cx.stop;
    return p;
}
struct TmpFrame {
    Ctx &cx;
    uint32_t mark;
    PD_HD explicit TmpFrame(Ctx &c) : cx(c), mark(c.stop) {}
    PD_HD ~TmpFrame() { cx.stop = mark; }
};

// ---------------------------------------------------------------------------
// a trick in progress (bridge.cl:1535-1620)
// ---------------------------------------------------------------------------
struct Trick {
    int8_t suit, trump, hsuit, hrank, winner;   // suit / trump / hsuit: -1 = nil
    uint8_t n;
    Card cards[4];
    Hand rem[4];
};
PD_HD void trick_init(Trick &t, int suit, int trump)
{
    t.suit = (int8_t)suit; t.trump = (int8_t)trump; t.hsuit = (int8_t)suit; t.hrank = -1; t.winner = -1; t.n = 0;
}
// card-higher (1556-1570)
PD_HD bool card_higher(const Trick &t, int asuit, int arank, int bsuit, int brank)
{
    const bool atr = asuit == t.trump, as = asuit == t.suit, btr = bsuit == t.trump, bs = bsuit == t.suit;
    if (atr && !btr) return true;
    if (btr && !atr) return false;
    if (as && !bs) return true;
    if (bs && !as) return false;
    if (!(as || atr)) return false;
    if (!(bs || btr)) return true;
    return arank > brank;
}
// (play trick card rest) 1588-1599
PD_HDN void trick_play(Trick &t, Card card, const Hand &rest)
{
    if (card != NOCARD && card_higher(t, csuit(card), crank(card), t.hsuit, t.hrank)) {
        t.hsuit = (int8_t)csuit(card); t.hrank = (int8_t)crank(card); t.winner = (int8_t)t.n;
    }
    t.rem[t.n] = rest;
    t.cards[t.n] = card;
    ++t.n;
}
PD_HD bool winnerp(const Trick &t) { return t.winner >= 0 && (t.n & 1) == (t.winner & 1); }

// (play hand suit chooser): the card of `rank` leaves the hand
PD_HD void hand_take(const Hand &h, int suit, int rank, Card &card, Hand &rest)
{
    rest = h;
    rest.s.m = h.s.m & ~(1ull << (16 * suit + rank));
    card = mk(suit, rank);
}

// weak-suit (1732-1768): the suit to discard from; -1 when the hand is empty
PD_HDN int weak_suit(const Hand &h, const Trick &t)
{
    int acc = -1;
    PD_ROLLED
    for (int suit = 0; suit < 4; ++suit) {
        const uint32_t cards = h.s[suit];
        if (!cards) continue;
        if (acc < 0) { acc = suit; continue; }
        const uint32_t accards = h.s[acc];
        if (suit == t.trump) continue;
        const bool hon = (cards >> 9) != 0, acchon = (accards >> 9) != 0;      // a card above rank 8
        if (acc == t.trump) { acc = suit; continue; }
        if (!hon && !acchon) { if (popc(cards) < popc(accards)) acc = suit; continue; }
        if (!hon) { acc = suit; continue; }
        if (!acchon) continue;
        int na = 0, nb = 0;
        PD_ROLLED
        for (int k = 0; k < t.n; ++k) {
            na += popc(above(t.rem[k].s[suit], highbit(cards)));
            nb += popc(above(t.rem[k].s[acc], highbit(accards)));
        }
        const int buf_a = popc(cards) - na, buf_b = popc(accards) - nb;
        if (buf_a == buf_b) { if (lowbit(cards) < lowbit(accards)) acc = suit; }
        else if ((buf_a > 0 && buf_b > 0) || (buf_a < 0 && buf_b < 0)) { if (buf_a > buf_b) acc = suit; }
        else if (buf_a > 0) acc = suit;
    }
    return acc;
}

// chooser results: rank to play or -1
PD_HD int ch_lowest(uint32_t m) { return m ? lowbit(m) : -1; }
PD_HD int ch_first_above(uint32_t m, int r) { const uint32_t a = above(m, r); return a ? lowbit(a) : -1; }
// take-higher (1518-1526)
PD_HD int ch_take_higher(uint32_t m, int high, bool use_highest)
{
    if (use_highest) { const int h = highbit(m); return h > high ? h : -1; }
    return ch_first_above(m, high);
}
// closest-card (1787-1790): first card >= rank, else the highest
PD_HD int ch_closest(uint32_t m, int rank) { const uint32_t a = above(m, rank - 1); return a ? lowbit(a) : highbit(m); }

struct Play { bool ok; Card card; Hand rest; };
PD_HDN Play play_rank(const Hand &h, int suit, int rank)
{
    Play p; p.ok = rank >= 0;
    if (p.ok) hand_take(h, suit, rank, p.card, p.rest); else { p.card = NOCARD; p.rest = h; }
    return p;
}
PD_HDN Play play_lowest(Ctx &cx, const Hand &h, int suit)
{
    if (suit < 0) { fail(cx, PD_E_LISP); Play p; p.ok = false; p.card = NOCARD; p.rest = h; return p; }   // (nth nil suits)
    return play_rank(h, suit, ch_lowest(h.s[suit]));
}

// beat-or-low (1804-1834)
PD_HDN bool beat_or_low(Ctx &cx, Trick &t, const Hand &h, bool use_highest)
{
    const int suit = t.suit, trump = t.trump, hsuit = t.hsuit, hrank = t.hrank;
    Play p; p.ok = false;
    if (hsuit >= 0 && hsuit == trump && suit != trump) {
        p = play_lowest(cx, h, suit);
        if (!p.ok && !winnerp(t)) p = play_rank(h, trump, ch_first_above(h.s[trump], hrank));
    } else if (hsuit >= 0 && hsuit == suit) {
        const uint32_t m = h.s[suit];
        if (m) {
            int r = -1;
            if (!winnerp(t)) r = ch_take_higher(m, hrank, use_highest);
            if (r < 0) r = lowbit(m);
            p = play_rank(h, suit, r);
        }
        if (!p.ok && trump >= 0) p = play_lowest(cx, h, trump);
    } else {
        if (suit < 0) { fail(cx, PD_E_LISP); return false; }
        const uint32_t m = h.s[suit];
        if (m) {
            int r = ch_take_higher(m, -1, use_highest);
            if (r < 0) r = lowbit(m);
            p = play_rank(h, suit, r);
        }
        if (!p.ok && trump >= 0) p = play_lowest(cx, h, trump);
    }
    if (!p.ok) {
        p = play_lowest(cx, h, weak_suit(h, t));
        if (!p.ok && trump >= 0) p = play_lowest(cx, h, trump);
    }
    if (!p.ok) return false;
    trick_play(t, p.card, p.rest);
    return true;
}

// finesse-if-possible (1905-1947)
PD_HDN bool finesse_if_possible(Ctx &cx, Trick &t, const Hand &h1, const Hand &h2)
{
    const int suit = t.suit, trump = t.trump;
    if (suit < 0) { fail(cx, PD_E_LISP); return false; }
    const bool wp = winnerp(t);
    const int hsuit = wp ? -1 : t.hsuit, hrank = wp ? -1 : t.hrank;
    const uint32_t h2suit = h2.s[suit];
    Play p; p.ok = false;
    bool trump_clause = false;
    if (hsuit >= 0 && hsuit == trump && trump != suit) {
        p = play_lowest(cx, h1, suit);
        if (!p.ok) {
            if (h2suit && suit != trump) p = play_rank(h1, trump, ch_first_above(h1.s[trump], hrank));
            else {
                const int h2max = h2.s[trump] ? highbit(h2.s[trump]) : -1;
                p = play_rank(h1, trump, ch_first_above(h1.s[trump], hrank > h2max ? hrank : h2max));
            }
        }
    } else if (hsuit >= 0 && hsuit == suit) {
        const uint32_t m = h1.s[suit];
        if (m) {
            int r = -1;
            if (h2suit) {
                const int h2max = highbit(h2suit);
                r = ch_first_above(m, h2max > hrank ? h2max : hrank);
            }
            if (r < 0) r = ch_first_above(m, hrank);
            if (r < 0) r = lowbit(m);
            p = play_rank(h1, suit, r);
        }
        trump_clause = true;
    } else {
        const uint32_t m = h1.s[suit];
        if (m) {
            int r = -1;
            if (h2suit) r = ch_first_above(m, highbit(h2suit));
            if (r < 0) r = lowbit(m);
            p = play_rank(h1, suit, r);
        }
        trump_clause = true;
    }
    if (!p.ok && trump_clause) {
        if (trump >= 0 && !h2suit && h2.s[trump]) p = play_rank(h1, trump, ch_first_above(h1.s[trump], highbit(h2.s[trump])));
        else if (trump >= 0 && !wp) p = play_lowest(cx, h1, trump);
    }
    if (!p.ok) p = play_lowest(cx, h1, weak_suit(h1, t));
    if (!p.ok) return false;
    trick_play(t, p.card, p.rest);
    return true;
}

// beats? (1705-1714)
PD_HDN bool beats(const Hand &a, const Hand &b, int suit, int trump)
{
    const uint32_t ac = a.s[suit], bc = b.s[suit];
    const uint32_t atr = trump >= 0 ? a.s[trump] : 0u, btr = trump >= 0 ? b.s[trump] : 0u;
    if (ac && bc) return highbit(ac) > highbit(bc);
    if (!ac && !bc && atr && btr) return highbit(atr) > highbit(btr);
    if (!ac && atr) return true;
    return false;
}

// ---------------------------------------------------------------------------
// outcomes (bridge.cl:2204-2267)
// ---------------------------------------------------------------------------
// struct copies go through one out-of-line routine each: the search is some hundred KB of code with no hot loop, and
// instruction fetch is a quarter of its time (profiles/r02_prof_playdeal.txt)
PD_HDN void copy_words(void *dst, const void *src, int words)
{
    uint32_t *d = static_cast<uint32_t *>(dst);
    const uint32_t *s = static_cast<const uint32_t *>(src);
    PD_ROLLED
    for (int i = 0; i < words; ++i) d[i] = s[i];
}
PD_HDN void outcome_copy(Outcome &dst, const Outcome &src)
{
    static_assert(sizeof(Outcome) % 16 == 0, "Outcome is copied in 16-byte pieces");
    if (&dst == &src) return;
    struct alignas(16) Q { uint32_t w[4]; };
    Q *d = reinterpret_cast<Q *>(&dst);
    const Q *q = reinterpret_cast<const Q *>(&src);
    for (int i = 0; i < (int)(sizeof(Outcome) / 16); ++i) d[i] = q[i];
}

PD_HDN void outcome_empty(Outcome &o, const Hand *hands)
{
    o.n = 0; o.roll = 0; o.score[0] = o.score[1] = 0; o.valid = 1;
    PD_ROLLED
    for (int i = 0; i < 4; ++i) o.rem[i] = hands[i];
}
PD_HD void outcome_nil(Outcome &o) { o.valid = 0; o.n = 0; o.roll = 0; o.score[0] = o.score[1] = 0; }

// new-outcome of a finished trick (2204-2210)
PD_HDN void outcome_of_trick(Ctx &cx, Outcome &o, const Trick &t)
{
    if (t.winner < 0 || t.n != 4) { fail(cx, PD_E_LISP); outcome_nil(o); return; }
    o.valid = 1; o.n = 1; o.roll = (uint8_t)t.winner;
    o.score[0] = (t.winner & 1) ? 0 : 1; o.score[1] = (t.winner & 1) ? 1 : 0;
    o.tricks[0] = pack4(t.cards);
    o.wno[0] = (uint8_t)t.winner;
    PD_ROLLED
    for (int i = 0; i < 4; ++i) o.rem[i] = t.rem[(i + t.winner) & 3];       // (roll (- winner) remaining)
}
// add-outcome (2212-2223): self += other, roll := other's roll.  star = add-outcome* (2225-2235): roll adds up.
PD_HDN void outcome_add(Ctx &cx, Outcome &self, const Outcome &other, bool star)
{
    if (self.n + other.n > 13) { fail(cx, PD_E_CAP); return; }
    PD_ROLLED
    for (int k = 0; k < other.n; ++k) {
        self.tricks[self.n + k] = other.tricks[k];
        self.wno[self.n + k] = other.wno[k];
    }
    const bool flip = (self.roll & 1) != 0;
    self.score[0] = (uint8_t)(self.score[0] + other.score[flip ? 1 : 0]);
    self.score[1] = (uint8_t)(self.score[1] + other.score[flip ? 0 : 1]);
    self.n = (uint8_t)(self.n + other.n);
    self.roll = star ? (uint8_t)((self.roll + other.roll) & 3) : other.roll;
    PD_ROLLED
    for (int i = 0; i < 4; ++i) self.rem[i] = other.rem[i];
}
PD_HD bool outcome_won(const Outcome &o) { return o.n > 0 && (o.wno[0] & 1) == 0; }

// simple-trick (2269-2277).  This is the estimator's inner loop (18 000 calls per deal): the trick is kept in REGISTERS
// (`TrickR`) and the four plays run through one rolled loop that holds a single inlined copy of beat-or-low and of
// finesse-if-possible -- a third of the instructions and half of the code of four calls on a `Trick` in memory, which
// is what the colder callers (hunt-card, reproduce) still use.  Same decisions, statement by statement.
struct TrickR {
    int suit, trump, hsuit, hrank, winner, n;
    uint32_t cards;              // card i in byte i
    uint64_t rem0, rem1, rem2, rem3;    // the hands after each play: scalars, so that they stay in registers
};
PD_HD uint32_t hsuit_of(uint64_t h, int suit) { return (uint32_t)(h >> (16 * suit)) & 0x1FFFu; }
PD_HD bool winnerp_r(const TrickR &t) { return t.winner >= 0 && (t.n & 1) == (t.winner & 1); }
PD_HD void play_r(TrickR &t, int psuit, int prank, uint64_t hand)
{
    const bool atr = psuit == t.trump, as = psuit == t.suit, btr = t.hsuit == t.trump, bs = t.hsuit == t.suit;
    bool higher;                                                     // card-higher (1556-1570)
    if (atr != btr) higher = atr;
    else if (as != bs) higher = as;
    else if (!(as || atr)) higher = false;
    else if (!(bs || btr)) higher = true;
    else higher = prank > t.hrank;
    if (higher) { t.hsuit = psuit; t.hrank = prank; t.winner = t.n; }
    const uint64_t rest = hand & ~(1ull << (16 * psuit + prank));
    t.rem0 = t.n == 0 ? rest : t.rem0; t.rem1 = t.n == 1 ? rest : t.rem1;
    t.rem2 = t.n == 2 ? rest : t.rem2; t.rem3 = t.n == 3 ? rest : t.rem3;
    t.cards |= (uint32_t)mk(psuit, prank) << (8 * t.n);
    ++t.n;
}
// weak-suit (1732-1768) on a hand word
PD_HD int weak_suit_r(uint64_t h, const TrickR &t)
{
    int acc = -1;
    PD_ROLLED
    for (int suit = 0; suit < 4; ++suit) {
        const uint32_t cards = hsuit_of(h, suit);
        if (!cards) continue;
        if (acc < 0) { acc = suit; continue; }
        const uint32_t accards = hsuit_of(h, acc);
        if (suit == t.trump) continue;
        const bool hon = (cards >> 9) != 0, acchon = (accards >> 9) != 0;
        if (acc == t.trump) { acc = suit; continue; }
        if (!hon && !acchon) { if (popc(cards) < popc(accards)) acc = suit; continue; }
        if (!hon) { acc = suit; continue; }
        if (!acchon) continue;
        int na = 0, nb = 0;
        const int ha = highbit(cards), hb = highbit(accards);
        if (t.n > 0) { na += popc(above(hsuit_of(t.rem0, suit), ha)); nb += popc(above(hsuit_of(t.rem0, acc), hb)); }
        if (t.n > 1) { na += popc(above(hsuit_of(t.rem1, suit), ha)); nb += popc(above(hsuit_of(t.rem1, acc), hb)); }
        if (t.n > 2) { na += popc(above(hsuit_of(t.rem2, suit), ha)); nb += popc(above(hsuit_of(t.rem2, acc), hb)); }
        if (t.n > 3) { na += popc(above(hsuit_of(t.rem3, suit), ha)); nb += popc(above(hsuit_of(t.rem3, acc), hb)); }
        const int buf_a = popc(cards) - na, buf_b = popc(accards) - nb;
        if (buf_a == buf_b) { if (lowbit(cards) < lowbit(accards)) acc = suit; }
        else if ((buf_a > 0 && buf_b > 0) || (buf_a < 0 && buf_b < 0)) { if (buf_a > buf_b) acc = suit; }
        else if (buf_a > 0) acc = suit;
    }
    return acc;
}
// the discard both routines end with: lowest card of the weak suit ((nth nil suits) is a Lisp error)
PD_HD void discard_r(Ctx &cx, const TrickR &t, uint64_t h, int &ps, int &pr)
{
    const int ws = weak_suit_r(h, t);
    if (ws < 0) { fail(cx, PD_E_LISP); return; }
    const uint32_t m = hsuit_of(h, ws);
    if (m) { ps = ws; pr = lowbit(m); }
}
PD_HD void lowest_r(uint64_t h, int suit, int &ps, int &pr)
{
    const uint32_t m = hsuit_of(h, suit);
    if (m) { ps = suit; pr = lowbit(m); }
}
// beat-or-low (1804-1834) up to its last resort, the discard: which card; -1 = none
PD_HD void beat_or_low_r(Ctx &cx, const TrickR &t, uint64_t h, bool use_highest, int &ps, int &pr)
{
    const int suit = t.suit, trump = t.trump, hsuit = t.hsuit, hrank = t.hrank;
    const bool wp = winnerp_r(t);
    ps = pr = -1;
    if (hsuit >= 0 && hsuit == trump && suit != trump) {
        if (suit < 0) fail(cx, PD_E_LISP); else lowest_r(h, suit, ps, pr);
        if (pr < 0 && !wp) { const int r = ch_first_above(hsuit_of(h, trump), hrank); if (r >= 0) { ps = trump; pr = r; } }
    } else {
        const bool follow = hsuit >= 0 && hsuit == suit;
        if (!follow && suit < 0) { fail(cx, PD_E_LISP); return; }
        const uint32_t m = hsuit_of(h, suit);
        if (m) {
            int r = -1;
            if (!follow || !wp) r = ch_take_higher(m, follow ? hrank : -1, use_highest);
            ps = suit; pr = r < 0 ? lowbit(m) : r;
        }
        if (pr < 0 && trump >= 0) lowest_r(h, trump, ps, pr);
    }
    // the caller discards when nothing was picked (and then tries the lowest trump)
}
// finesse-if-possible (1905-1947) up to its last resort, the discard: which card; -1 = none
PD_HD void finesse_r(Ctx &cx, const TrickR &t, uint64_t h1, uint64_t h2, int &ps, int &pr)
{
    const int suit = t.suit, trump = t.trump;
    ps = pr = -1;
    if (suit < 0) { fail(cx, PD_E_LISP); return; }
    const bool wp = winnerp_r(t);
    const int hsuit = wp ? -1 : t.hsuit, hrank = wp ? -1 : t.hrank;
    const uint32_t h2suit = hsuit_of(h2, suit);
    bool trump_clause = false;
    if (hsuit >= 0 && hsuit == trump && trump != suit) {
        lowest_r(h1, suit, ps, pr);
        if (pr < 0) {
            int over = hrank;
            if (!(h2suit && suit != trump)) { const uint32_t h2t = hsuit_of(h2, trump); const int h2max = h2t ? highbit(h2t) : -1; if (h2max > over) over = h2max; }
            const int r = ch_first_above(hsuit_of(h1, trump), over);
            if (r >= 0) { ps = trump; pr = r; }
        }
    } else {
        const bool follow = hsuit >= 0 && hsuit == suit;
        const uint32_t m = hsuit_of(h1, suit);
        if (m) {
            int r = -1;
            if (h2suit) { const int h2max = highbit(h2suit); r = ch_first_above(m, follow && hrank > h2max ? hrank : h2max); }
            if (r < 0 && follow) r = ch_first_above(m, hrank);
            ps = suit; pr = r < 0 ? lowbit(m) : r;
        }
        trump_clause = true;
    }
    if (pr < 0 && trump_clause && trump >= 0) {
        const uint32_t h2t = hsuit_of(h2, trump);
        if (!h2suit && h2t) { const int r = ch_first_above(hsuit_of(h1, trump), highbit(h2t)); if (r >= 0) { ps = trump; pr = r; } }
        else if (!wp) lowest_r(h1, trump, ps, pr);
    }
    // the caller discards when nothing was picked
}

PD_HDN void simple_trick(Ctx &cx, Outcome &o, int suit, const Hand *h, bool use_highest)
{
    if (h[0].s[suit] == 0) { outcome_empty(o, h); return; }
    TrickR t;
    t.suit = suit; t.trump = cx.trump; t.hsuit = suit; t.hrank = -1; t.winner = -1; t.n = 0; t.cards = 0;
    t.rem0 = t.rem1 = t.rem2 = t.rem3 = 0;
    const bool cover = beats(h[3], h[2], suit, cx.trump);            // second hand: beat-or-low instead of a finesse
    PD_ROLLED
    for (int pos = 0; pos < 4; ++pos) {
        const uint64_t hand = h[pos].s.m;
        int ps, pr;
        const bool fin = pos == 2 || (pos == 1 && !cover);
        if (fin) finesse_r(cx, t, hand, h[(pos + 1) & 3].s.m, ps, pr);
        else beat_or_low_r(cx, t, hand, pos == 0 && use_highest, ps, pr);
        if (pr < 0 && !cx.status) {                                  // both end with a discard from the weak suit
            discard_r(cx, t, hand, ps, pr);
            if (pr < 0 && !fin && t.trump >= 0) lowest_r(hand, t.trump, ps, pr);
        }
        if (pr >= 0) play_r(t, ps, pr, hand);
    }
    // new-outcome of the finished trick (2204-2210)
    if (t.winner < 0 || t.n != 4) { fail(cx, PD_E_LISP); outcome_nil(o); return; }
    o.valid = 1; o.n = 1; o.roll = (uint8_t)t.winner;
    o.score[0] = (t.winner & 1) ? 0 : 1; o.score[1] = (t.winner & 1) ? 1 : 0;
    o.tricks[0] = t.cards;
    o.wno[0] = (uint8_t)t.winner;
    {                                                                // (roll (- winner) remaining)
        const bool by2 = (t.winner & 2) != 0, by1 = (t.winner & 1) != 0;
        const uint64_t a0 = by2 ? t.rem2 : t.rem0, a1 = by2 ? t.rem3 : t.rem1, a2 = by2 ? t.rem0 : t.rem2, a3 = by2 ? t.rem1 : t.rem3;
        o.rem[0].s.m = by1 ? a1 : a0; o.rem[1].s.m = by1 ? a2 : a1; o.rem[2].s.m = by1 ? a3 : a2; o.rem[3].s.m = by1 ? a0 : a3;
    }
}

// ---------------------------------------------------------------------------
// the three look-ahead searches: suit-tricks' play-through (2280-2288), play-routes' suit-tops (2477-2483),
// immediate-tricks (2705-2722).  All are "try a list of (suit, lead high?, which partner leads) options, follow each
// trick that qualifies with the same search on the hands it leaves, keep the best by the leader's side's tricks".
// ---------------------------------------------------------------------------
enum SearchKind { PLAY_THROUGH = 0, SUIT_TOPS = 1, IMMEDIATE = 2 };

struct SearchOpt { int8_t suit; uint8_t top, rolled; };
struct SearchFrame {
    Hand hands[4];
    SearchOpt opts[12];
    uint8_t nopt, i;
    Outcome best;        // valid = 0: nil
    uint8_t have_best;   // PLAY_THROUGH / SUIT_TOPS: the first candidate initialises the accumulator
    Outcome cur;         // the trick whose continuation is being searched
};

PD_HD void roll2(const Hand *h, Hand *out) { out[0] = h[2]; out[1] = h[3]; out[2] = h[0]; out[3] = h[1]; }

// candidate list of immediate-tricks for the hands of a frame
PD_HDN void immediate_opts(const Ctx &cx, SearchFrame &f, const int8_t *suits, int nsuits)
{
    f.nopt = 0;
    const Hand &a = f.hands[0], &b = f.hands[1], &d = f.hands[3];
    PD_ROLLED
    for (int k = 0; k < nsuits; ++k) {
        const int suit = suits[k];
        const uint32_t cards = a.s[suit];
        int n = 0;
        if (suit == cx.trump) {
            if (cards && (b.s[suit] || d.s[suit])) n = popc(cards) > 1 ? 2 : 1;
        } else if (cards) n = popc(cards) < 2 ? 1 : 2;
        PD_ROLLED
        for (int j = 0; j < n && f.nopt < 12; ++j) { f.opts[f.nopt].suit = (int8_t)suit; f.opts[f.nopt].top = (uint8_t)j; f.opts[f.nopt].rolled = 0; ++f.nopt; }
    }
}
PD_HD void fourway_opts(SearchFrame &f, int suit)
{
    f.nopt = 4;
    PD_ROLLED
    for (int j = 0; j < 4; ++j) { f.opts[j].suit = (int8_t)suit; f.opts[j].top = (uint8_t)(j < 2); f.opts[j].rolled = (uint8_t)(j & 1); }
}

// acc := (best-outcome acc val) (2253-2258)
PD_HDN void best_update(SearchFrame &f, const Outcome &val)
{
    if (!f.have_best) { outcome_copy(f.best, val); f.have_best = 1; return; }
    if (!f.best.valid || (val.valid && f.best.score[0] < val.score[0])) outcome_copy(f.best, val);
}

// The two exhaustive look-aheads (PLAY_THROUGH, SUIT_TOPS) reach the same four hands along many orders of play -- ten
// frames for every distinct position, and with them more than half of all simple-tricks of a deal -- and the value of a
// frame is a function of its hands alone (same suit, same trump, no state outside the frame).  Each such search
// therefore keeps the values of the positions it has finished, keyed by the four hand words, in the arena: a chained
// hash table that lives as long as the call.
struct MemoEntry { uint64_t key[4]; uint32_t next; uint32_t pad[3]; Outcome val; };
constexpr uint32_t MEMO_BUCKETS = 256;
PD_HD uint32_t memo_bucket(const Hand *h)
{
    uint64_t x = h[0].s.m ^ (h[1].s.m << 13 | h[1].s.m >> 51) ^ (h[2].s.m << 29 | h[2].s.m >> 35) ^ (h[3].s.m << 43 | h[3].s.m >> 21);
    x *= 0x9E3779B97F4A7C15ull;
    return (uint32_t)(x >> 56) & (MEMO_BUCKETS - 1u);
}
PD_HDN const Outcome *memo_find(Ctx &cx, uint32_t buckets, const Hand *h)
{
    uint32_t e = at<uint32_t>(cx, buckets)[memo_bucket(h)];
    PD_ROLLED
    while (e) {
        const MemoEntry *m = at<MemoEntry>(cx, e);
        if (m->key[0] == h[0].s.m && m->key[1] == h[1].s.m && m->key[2] == h[2].s.m && m->key[3] == h[3].s.m) return &m->val;
        e = m->next;
    }
    return nullptr;
}
PD_HDN void memo_put(Ctx &cx, uint32_t buckets, const Hand *h, const Outcome &val)
{
    const uint32_t pad = (16u - (cx.top & 15u)) & 15u;
    const uint32_t raw = arena_alloc(cx, (uint32_t)sizeof(MemoEntry) + pad);
    if (cx.status) return;
    const uint32_t off = raw + pad;
    MemoEntry *m = at<MemoEntry>(cx, off);
    uint32_t *head = at<uint32_t>(cx, buckets) + memo_bucket(h);
    PD_ROLLED
    for (int i = 0; i < 4; ++i) m->key[i] = h[i].s.m;
    m->next = *head;
    outcome_copy(m->val, val);
    PD_SYNC();
    *head = off;
    PD_SYNC();
}

// `stack`: MAXD frames in the caller's local memory
PD_HDN void search(Ctx &cx, SearchKind kind, int suit, const int8_t *suits, int nsuits, const Hand *hands, SearchFrame *stack, Outcome &result)
{
    TmpFrame tf(cx);
    Hand *hs = tmp<Hand>(cx, 4);
    Outcome &cand = *tmp<Outcome>(cx);
    int sp = 0;
    const bool memo = kind != IMMEDIATE;
    const uint32_t mark = cx.top;
    uint32_t buckets = 0;
    if (memo) {
        buckets = arena_alloc(cx, MEMO_BUCKETS * 4u);
        if (cx.status) { outcome_nil(result); return; }
        uint32_t *b = at<uint32_t>(cx, buckets);
        PD_ROLLED
        for (uint32_t i = PD_LANE; i < MEMO_BUCKETS; i += PD_LANES) b[i] = 0u;
        PD_SYNC();
    }
    auto open = [&](SearchFrame &f, const Hand *h) {
        PD_ROLLED
        for (int i = 0; i < 4; ++i) f.hands[i] = h[i];
        f.i = 0;
        if (kind == IMMEDIATE) { immediate_opts(cx, f, suits, nsuits); outcome_empty(f.best, h); f.have_best = 1; }
        else { fourway_opts(f, suit); f.have_best = 0; outcome_nil(f.best); }
    };
    open(stack[0], hands);
    for (;;) {
        SearchFrame &f = stack[sp];
        if (cx.status) { outcome_nil(result); return; }
        if (f.i == f.nopt) {
            // the frame's value
            Outcome &val = f.best;
            if (kind == SUIT_TOPS && !val.valid) outcome_empty(val, f.hands);
            if (kind == PLAY_THROUGH && !f.have_best) outcome_empty(val, f.hands);
            if (sp == 0) { outcome_copy(result, val); cx.top = mark; return; }
            if (memo) memo_put(cx, buckets, f.hands, val);
            SearchFrame &p = stack[sp - 1];
            outcome_add(cx, p.cur, val, false);          // (add-outcome self (apply if-won remaining))
            best_update(p, p.cur);
            --sp;
            continue;
        }
        const SearchOpt o = f.opts[f.i++];
        if (o.rolled) roll2(f.hands, hs); else for (int i = 0; i < 4; ++i) hs[i] = f.hands[i];
        simple_trick(cx, f.cur, o.suit, hs, o.top != 0);
        const bool follow = f.cur.n > 0 && (kind == PLAY_THROUGH || outcome_won(f.cur));
        if (follow) {
            if (memo) {
                const Outcome *known = memo_find(cx, buckets, f.cur.rem);
                if (known) { outcome_add(cx, f.cur, *known, false); best_update(f, f.cur); continue; }
            }
            if (sp + 1 >= MAXD) { fail(cx, PD_E_CAP); outcome_nil(result); return; }
            ++sp;
            open(stack[sp], f.cur.rem);
            continue;
        }
        // do-next gave nil
        if (kind == PLAY_THROUGH) outcome_empty(cand, hs); else outcome_nil(cand);
        best_update(f, cand);
    }
}

// ---------------------------------------------------------------------------
// suit-tricks -> suit-value -> deal-plan -> find-opp-weakness (2279-2356, 2745-2759)
// ---------------------------------------------------------------------------
struct alignas(4) SuitValue {
    int8_t suit, trump;          // trump: -1 = evaluated without
    uint8_t ngroups;
    uint8_t glen[15];
    Card gcards[15][13];
    int ours, theirs;
};

PD_HDN void suit_value(Ctx &cx, SuitValue &sv, int suit, int trump, const Hand *deal, SearchFrame *stack)
{
    TmpFrame tf(cx);
    const int saved = cx.trump;
    cx.trump = trump;
    Outcome &o = *tmp<Outcome>(cx);
    search(cx, PLAY_THROUGH, suit, nullptr, 0, deal, stack, o);
    cx.trump = saved;
    sv.suit = (int8_t)suit; sv.trump = (int8_t)trump; sv.ngroups = 0;
    PD_ROLLED
    for (int i = 0; i < 15; ++i) sv.glen[i] = 0;
    // the reference folds the (leader's side won?, winning card) pairs from the last trick backwards (2289-2296);
    // forwards that is: a lost trick opens the next group, every trick's winning card joins the current group
    int g = 0;
    bool any = false;
    PD_ROLLED
    for (int k = 0; k < o.n && !cx.status; ++k) {
        const bool first = (o.wno[k] & 1) == 0;
        if (!first) ++g;
        if (g >= 15) { fail(cx, PD_E_CAP); break; }
        sv.gcards[g][sv.glen[g]++] = tcard(o.tricks[k], o.wno[k]);
        any = true;
    }
    sv.ngroups = any ? (uint8_t)(g + 1) : 0;
    sv.ours = sv.theirs = 0;
    PD_ROLLED
    for (int i = 0; i < sv.ngroups; ++i) { if (i & 1) sv.theirs += sv.glen[i]; else sv.ours += sv.glen[i]; }
}

struct Target {                  // one row of find-opp-weakness: (suit, ours - theirs, our ranks, their ranks)
    int8_t suit;
    int8_t diff;
    uint8_t nours, ntheirs, head;    // head: ranks of `theirs` already popped by play-strategy (2863)
    uint8_t theirs[13];
};
struct RootStrategy {            // soonest / establish of a play-strategy: shared by every strategy derived from it
    uint8_t nsoon, nest;
    int8_t soonest[MAXT];
    Target establish[MAXT];
};

PD_HDN void find_opp_weakness(Ctx &cx, RootStrategy &rs, const Hand *deal, SearchFrame *stack)
{
    const int trump = cx.trump;
    TmpFrame tf(cx);
    SuitValue *vals = tmp<SuitValue>(cx, 4);
    SuitValue &nt = *tmp<SuitValue>(cx), &tr = *tmp<SuitValue>(cx);
    int nv = 0;
    PD_ROLLED
    for (int suit = 0; suit < 4 && !cx.status; ++suit) {
        suit_value(cx, nt, suit, -1, deal, stack);
        if (trump >= 0) {
            suit_value(cx, tr, suit, trump, deal, stack);
            if (tr.ours > nt.ours) copy_words(&nt, &tr, (int)(sizeof(SuitValue) / 4));                 // better (2326-2327)
        }
        if (nt.ours > 0) copy_words(&vals[nv++], &nt, (int)(sizeof(SuitValue) / 4));
    }
    // (sort ... (> (ours a) (ours b))): stable
    PD_ROLLED
    for (int i = 1; i < nv; ++i) {
        PD_ROLLED
        for (int j = i; j > 0 && vals[j].ours > vals[j - 1].ours; --j) { copy_words(&nt, &vals[j], (int)(sizeof(SuitValue) / 4)); copy_words(&vals[j], &vals[j - 1], (int)(sizeof(SuitValue) / 4)); copy_words(&vals[j - 1], &nt, (int)(sizeof(SuitValue) / 4)); }
    }
    int *order = tmp<int>(cx, 12), no = 0;
    if (trump >= 0) {
        PD_ROLLED
        for (int i = 0; i < nv; ++i) if (vals[i].trump >= 0) order[no++] = i;
        PD_ROLLED
        for (int i = 0; i < nv; ++i) if (vals[i].suit == trump) order[no++] = i;
        PD_ROLLED
        for (int i = 0; i < nv; ++i) if (!(vals[i].suit == trump || vals[i].trump >= 0)) order[no++] = i;
    } else for (int i = 0; i < nv; ++i) order[no++] = i;
    rs.nsoon = rs.nest = 0;
    PD_ROLLED
    for (int k = 0; k < no; ++k) {
        const SuitValue &sv = vals[order[k]];
        Target &t = *tmp<Target>(cx);
        t.suit = sv.suit; t.diff = (int8_t)(sv.ours - sv.theirs); t.nours = t.ntheirs = t.head = 0;
        PD_ROLLED
        for (int g = 0; g < sv.ngroups; ++g) {
            PD_ROLLED
            for (int j = 0; j < sv.glen[g]; ++j) {
                const Card c = sv.gcards[g][j];
                if (c == NOCARD || csuit(c) != sv.suit) continue;
                if (g & 1) t.theirs[t.ntheirs++] = (uint8_t)crank(c); else ++t.nours;
            }
        }
        if (t.diff == 0 && t.nours == 0) { if (rs.nsoon < MAXT) rs.soonest[rs.nsoon++] = t.suit; else fail(cx, PD_E_CAP); }
        if (t.diff > 0) { if (rs.nest < MAXT) rs.establish[rs.nest++] = t; else fail(cx, PD_E_CAP); }
    }
}

// ---------------------------------------------------------------------------
// routes (bridge.cl:2358-2660)
// ---------------------------------------------------------------------------
struct Route {                   // play-route; lives in the arena, identity = its offset
    int8_t frm, to, tvoid;       // -1 = nil
    uint8_t pad;
    uint32_t tricks;             // arena offset of len packed tricks
    uint32_t len;
    uint32_t mlo, mhi;           // bit cidx(c) for every card c that leads or is third in a trick of the route -- or was,
};                               // before tricks were dropped (a superset): first-pass skips routes the cards played miss
PD_HD uint64_t route_mask(const Route &r) { return (uint64_t)r.mlo | ((uint64_t)r.mhi << 32); }
struct Com {                     // play-com
    uint8_t zero_id;
    uint8_t nact, npend, nguard;
    uint32_t act[MAXR], pend[MAXR];
    int8_t gsuit[MAXG];
    uint16_t granks[MAXG];
};
struct RouteList { int n; uint32_t r[MAXR]; };


PD_HD uint64_t trick_mask(uint32_t t) { return (1ull << cidx(tcard(t, 0))) | (1ull << cidx(tcard(t, 2))); }
PD_HDN uint64_t tricks_mask(const uint32_t *src, int n)
{
    uint64_t m = 0;
    PD_ROLLED
    for (int i = 0; i < n; ++i) m |= trick_mask(src[i]);
    return m;
}
PD_HDN uint32_t new_route(Ctx &cx, int frm, int to, int tvoid, uint32_t tricks, uint32_t len, uint64_t mask)
{
    const uint32_t off = arena_alloc(cx, sizeof(Route));
    Route *r = at<Route>(cx, off);
    r->frm = (int8_t)frm; r->to = (int8_t)to; r->tvoid = (int8_t)tvoid; r->pad = 0; r->tricks = tricks; r->len = len;
    r->mlo = (uint32_t)mask; r->mhi = (uint32_t)(mask >> 32);
    return off;
}
PD_HD void rl_push(Ctx &cx, RouteList &l, uint32_t r) { if (l.n < MAXR) l.r[l.n++] = r; else fail(cx, PD_E_CAP); }
PD_HDN void rl_push_front(Ctx &cx, RouteList &l, uint32_t r)
{
    if (l.n >= MAXR) { fail(cx, PD_E_CAP); return; }
    PD_ROLLED
    for (int i = l.n; i > 0; --i) l.r[i] = l.r[i - 1];
    l.r[0] = r; ++l.n;
}

// the lanes copy a trick list together, four loads in flight per lane (the lists sit in L2)
PD_HD void copy_tricks(uint32_t *dst, const uint32_t *src, uint32_t n)
{
    uint32_t i = (uint32_t)PD_LANE;
    PD_ROLLED
    for (; i + 3u * PD_LANES < n; i += 4u * PD_LANES) {
        const uint32_t a = src[i], b = src[i + PD_LANES], c = src[i + 2u * PD_LANES], d = src[i + 3u * PD_LANES];
        dst[i] = a; dst[i + PD_LANES] = b; dst[i + 2u * PD_LANES] = c; dst[i + 3u * PD_LANES] = d;
    }
    PD_ROLLED
    for (; i < n; i += PD_LANES) dst[i] = src[i];
}
// add-at (2421-2425): a new route whose tricks are the concatenation
PD_HDN uint32_t add_at(Ctx &cx, uint32_t route, uint32_t tricks2, uint32_t len2, uint64_t mask2)
{
    const Route r = *at<Route>(cx, route);
    const uint32_t len = r.len + len2;
    const uint32_t off = arena_alloc(cx, len * 4u);
    if (cx.status) return route;
    uint32_t *dst = at<uint32_t>(cx, off);
    const uint32_t *a = at<uint32_t>(cx, r.tricks), *b = at<uint32_t>(cx, tricks2);
    copy_tricks(dst, a, r.len);
    copy_tricks(dst + r.len, b, len2);
    PD_SYNC();
    return new_route(cx, r.frm, r.to, r.tvoid, off, len, route_mask(r) | mask2);
}
// insert-route (2553-2561): merge into the first route with the same from / to, else push in front
PD_HDN void insert_route(Ctx &cx, RouteList &routes, uint32_t route)
{
    const Route r = *at<Route>(cx, route);
    PD_ROLLED
    for (int p = 0; p < routes.n; ++p) {
        const Route *q = at<Route>(cx, routes.r[p]);
        if (q->frm == r.frm && q->to == r.to) { routes.r[p] = add_at(cx, routes.r[p], r.tricks, r.len, route_mask(r)); return; }
    }
    rl_push_front(cx, routes, route);
}

// reproduce (2369-2395): replay the route's next trick with the cards the hands hold now; a trick that cannot be
// won is dropped and the next one tried.  Pops the route IN PLACE.
PD_HDN void reproduce(Ctx &cx, uint32_t route, const Hand *hands, Outcome &ans)
{
    // The hands do not change while tricks are dropped, and a route is mostly copies of a few tricks (the pending lists
    // are merged with themselves at every trick): a trick that failed once in this call fails again -- remember the last
    // few and drop their copies without replaying them.
    constexpr int NTRIED = 24;
    TmpFrame tried_frame(cx);
    uint32_t *tried = tmp<uint32_t>(cx, NTRIED);
    int ntried = 0;
    for (;;) {
        Route *r = at<Route>(cx, route);
        if (r->len == 0 || cx.status) { outcome_nil(ans); return; }
#if defined(__CUDA_ARCH__)
        // 99.6 % of the pops drop a copy of a trick that has failed already: 32 tricks per step, one per lane, up to the
        // first one not seen yet (every lane pops the same count, so the route record stays identical across the lanes)
        uint32_t base;
        {
            const uint32_t len = r->len, list = r->tricks, mine = len > (uint32_t)PD_LANE ? at<uint32_t>(cx, list)[PD_LANE] : 0u;
            bool fresh = len > (uint32_t)PD_LANE;
            PD_ROLLED
            for (int i = 0; i < ntried && fresh; ++i) fresh = tried[i] != mine;
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, fresh);
            const uint32_t drop = bal ? (uint32_t)__ffs((int)bal) : (len < 32u ? len : 32u);     // tricks popped, the fresh one included
            base = __shfl_sync(0xFFFFFFFFu, mine, bal ? (int)drop - 1 : 0);
            r->tricks = list + 4u * drop; r->len = len - drop;           // every lane stores the same values
            __syncwarp();
            if (!bal) continue;
            if (ntried < NTRIED) tried[ntried++] = base;
        }
#else
        const uint32_t base = at<uint32_t>(cx, r->tricks)[0];
        r->tricks += 4; r->len -= 1;
        {
            bool seen = false;
            PD_ROLLED
            for (int i = 0; i < ntried && !seen; ++i) seen = tried[i] == base;
            if (seen) continue;
            if (ntried < NTRIED) tried[ntried++] = base;
        }
#endif
        Card ref[4];
        PD_ROLLED
        for (int i = 0; i < 4; ++i) ref[i] = tcard(base, i);
        if (ref[0] == NOCARD) { fail(cx, PD_E_LISP); outcome_nil(ans); return; }
        if (!((hands[0].s[csuit(ref[0])] >> crank(ref[0])) & 1u)) {          // the partner led it: (roll 2 base)
            const Card t0 = ref[0], t1 = ref[1]; ref[0] = ref[2]; ref[1] = ref[3]; ref[2] = t0; ref[3] = t1;
        }
        if (ref[0] == NOCARD) { fail(cx, PD_E_LISP); outcome_nil(ans); return; }
        const int suit = csuit(ref[0]);
        const uint32_t m0 = hands[0].s[suit];
        bool have = false;
        if (m0) {
            TmpFrame tf(cx);
            Trick &t = *tmp<Trick>(cx);
            trick_init(t, suit, cx.trump);
            Play p = play_rank(hands[0], suit, ch_closest(m0, crank(ref[0])));
            trick_play(t, p.card, p.rest);
            if (beats(hands[3], hands[2], suit, cx.trump)) beat_or_low(cx, t, hands[1], false);
            else finesse_if_possible(cx, t, hands[1], hands[2]);
            if (ref[2] == NOCARD) { fail(cx, PD_E_LISP); outcome_nil(ans); return; }
            const int s2 = csuit(ref[2]);
            Play q; q.ok = false;
            if (hands[2].s[s2]) q = play_rank(hands[2], s2, ch_closest(hands[2].s[s2], crank(ref[2])));
            if (!q.ok) q = play_lowest(cx, hands[2], weak_suit(hands[2], t));
            if (!q.ok) { fail(cx, PD_E_LISP); outcome_nil(ans); return; }
            trick_play(t, q.card, q.rest);
            beat_or_low(cx, t, hands[3], false);
            outcome_of_trick(cx, ans, t);
            have = true;
        }
        if (have && ans.valid && outcome_won(ans)) return;
    }
}

// mk-route (2453-2474)
PD_HDN void mk_route(Ctx &cx, uint32_t com_off, const Hand *hands, Outcome &acc)
{
    const Com &com = *at<Com>(cx, com_off);
    TmpFrame tf(cx);
    RouteList *pr = tmp<RouteList>(cx, 2);
    Outcome &ans = *tmp<Outcome>(cx);
    Hand *hs = tmp<Hand>(cx, 4);
    PD_ROLLED
    for (int w = 0; w < 2; ++w) {
        const int player = w * 2;
        pr[w].n = 0;
        // the reference's sort predicate (2458-2462) is not an ordering; under SBCL's merge sort it yields: routes with
        // from = to in reverse input order, then other routes with a `from` in reverse input order, then the rest
        PD_ROLLED
        for (int i = com.nact - 1; i >= 0; --i) { const Route *r = at<Route>(cx, com.act[i]); if (r->frm >= 0 && r->frm == player && r->frm == r->to) rl_push(cx, pr[w], com.act[i]); }
        PD_ROLLED
        for (int i = com.nact - 1; i >= 0; --i) { const Route *r = at<Route>(cx, com.act[i]); if (r->frm >= 0 && r->frm == player && r->frm != r->to) rl_push(cx, pr[w], com.act[i]); }
        PD_ROLLED
        for (int i = 0; i < com.nact; ++i) { const Route *r = at<Route>(cx, com.act[i]); if (r->frm < 0) rl_push(cx, pr[w], com.act[i]); }
    }
    outcome_empty(acc, hands);
    int cur = 0, head[2] = {0, 0};
    PD_ROLLED
    while (head[cur] < pr[cur].n && !cx.status) {
        PD_ROLLED
        for (int i = 0; i < 4; ++i) hs[i] = acc.rem[i];
        reproduce(cx, pr[cur].r[head[cur]], hs, ans);
        if (!ans.valid) { ++head[cur]; continue; }
        const bool same = ans.wno[0] == 0;
        outcome_add(cx, acc, ans, false);
        if (!same) cur ^= 1;
    }
}

// play-routes (2476-2551)
struct Item { Card cards[4]; uint8_t wno; uint8_t mine; uint8_t guard; int8_t gsuit; uint16_t granks; };

PD_HDN uint32_t play_routes(Ctx &cx, const Hand *h, const int8_t *suits, int nsuits, SearchFrame *stack);

PD_HDN uint32_t tricks_from(Ctx &cx, const uint32_t *src, int n)
{
    const uint32_t off = arena_alloc(cx, (uint32_t)n * 4u);
    if (cx.status) return 0;
    uint32_t *dst = at<uint32_t>(cx, off);
    PD_ROLLED
    for (int i = 0; i < n; ++i) dst[i] = src[i];
    return off;
}

PD_HDN uint32_t play_routes(Ctx &cx, const Hand *h, const int8_t *suits, int nsuits, SearchFrame *stack)
{
    // sections of the classification (2523-2530); each keeps REVERSE input order (divlist conses in front)
    enum { S_GUARD = 0, S_MY_LEFT, S_PARTNER_LEFT, S_MY_TRUMPED, S_PARTNER_TRUMPED, S_MY_LEAD, S_PARTNER_LEAD, S_MY, S_PARTNER, NSEC };
    TmpFrame tf(cx);
    uint32_t (*sec)[60] = reinterpret_cast<uint32_t (*)[60]>(tmp<uint32_t>(cx, NSEC * 60));
    int *nsec = tmp<int>(cx, NSEC);
    int8_t *gs = tmp<int8_t>(cx, 8);
    uint16_t *gr = tmp<uint16_t>(cx, 8);
    Outcome &tops = *tmp<Outcome>(cx);
    uint32_t *tmpt = tmp<uint32_t>(cx, 60);
    RouteList &routes = *tmp<RouteList>(cx);
    int ng = 0;
    PD_ROLLED
    for (int i = 0; i < NSEC; ++i) nsec[i] = 0;
    PD_ROLLED
    for (int k = 0; k < nsuits && !cx.status; ++k) {
        const int suit = suits[k];
        search(cx, SUIT_TOPS, suit, nullptr, 0, h, stack, tops);
        if (cx.status) break;
        PD_ROLLED
        for (int t = 0; t < tops.n; ++t) {
            Card c[4];
            c[0] = tcard(tops.tricks[t], 0); c[1] = tcard(tops.tricks[t], 1); c[2] = tcard(tops.tricks[t], 2); c[3] = tcard(tops.tricks[t], 3);
            const int w = tops.wno[t];
            const Card win = c[w];
            const bool mine = win != NOCARD && ((h[0].s[csuit(win)] >> crank(win)) & 1u);
            const bool left = (c[0] == NOCARD ? -1 : csuit(c[0])) != (c[2] == NOCARD ? -1 : csuit(c[2]));
            int s;
            if (w == 1) s = S_GUARD;
            else if (left && w == 0 && mine) s = S_MY_LEFT;
            else if (left && w == 0 && !mine) s = S_PARTNER_LEFT;
            else if (left && w == 2 && !mine) s = S_MY_TRUMPED;
            else if (left && w == 2 && mine) s = S_PARTNER_TRUMPED;
            else if (w == 2 && !mine) s = S_MY_LEAD;
            else if (w == 2 && mine) s = S_PARTNER_LEAD;
            else if (mine) s = S_MY;
            else s = S_PARTNER;
            if (s == S_GUARD) { fail(cx, PD_E_LISP); break; }      // a lost trick never reaches here (suit-tops follows won tricks only)
            if (nsec[s] >= 60) { fail(cx, PD_E_CAP); break; }
            sec[s][nsec[s]++] = pack4(c);
        }
        // guards (2484-2495): the top cards of the suit held by the hands to the right and left of the next leader,
        // down to the first card of the leader's side
        uint16_t ranks = 0;
        {
            const Hand *r = tops.rem;                       // (roll 1 remaining): hand ids 0..3 = r[3], r[0], r[1], r[2]
            const uint32_t opp = (uint32_t)r[3].s[suit] | r[1].s[suit], own = (uint32_t)r[0].s[suit] | r[2].s[suit];
            if (own) ranks = (uint16_t)above(opp, highbit(own));
            else ranks = (uint16_t)opp;
        }
        if (ranks) { if (ng < 8) { gs[ng] = (int8_t)suit; gr[ng] = ranks; ++ng; } else fail(cx, PD_E_CAP); }
    }
    if (cx.status) return 0;
    // routes in the order of 2542-2549, each section reversed
    routes.n = 0;
    auto section = [&](int s, int frm, int to) {
        PD_ROLLED
        for (int i = 0; i < nsec[s]; ++i) tmpt[i] = sec[s][nsec[s] - 1 - i];
        const uint32_t off = tricks_from(cx, tmpt, nsec[s]);
        rl_push(cx, routes, new_route(cx, frm, to, -1, off, (uint32_t)nsec[s], tricks_mask(tmpt, nsec[s])));
    };
    section(S_MY, -1, 0);
    section(S_PARTNER, -1, 2);
    section(S_MY_LEAD, 0, 2);
    section(S_PARTNER_LEAD, 2, 0);
    // void-vertices (2509-2518): the tricks are split by the suit LED (0, 1, 2, other) and the k-th split is paired with
    // the k-th element of `suits` -- the reference's pairing, kept as it is
    auto vertices = [&](int s, int frm, int to) {
        PD_ROLLED
        for (int k = 0; k < nsuits && k < 4; ++k) {
            int n = 0;
            PD_ROLLED
            for (int i = 0; i < nsec[s]; ++i) {
                const uint32_t t = sec[s][i];                    // input order; the split reverses it, like the section
                const Card c0 = tcard(t, 0);
                const int led = c0 == NOCARD ? -1 : csuit(c0);
                const int split = (led >= 0 && led <= 2) ? led : 3;
                if (split == k) tmpt[n++] = t;
            }
            if (n == 0) continue;
            // sec[s] is already in input order here; the section list handed to void-vertices is the reversed one and
            // divlist reverses it again: the route's tricks are in input order
            const uint32_t off = tricks_from(cx, tmpt, n);
            rl_push(cx, routes, new_route(cx, frm, to, suits[k], off, (uint32_t)n, tricks_mask(tmpt, n)));
        }
    };
    vertices(S_MY_LEFT, 0, 0);
    vertices(S_MY_TRUMPED, 0, 2);
    vertices(S_PARTNER_LEFT, 2, 2);
    vertices(S_PARTNER_TRUMPED, 2, 0);
    // mk-play-com (2439-2451): empty routes dropped; a route waits (pending) while both a and c hold its void suit
    const uint32_t coff = arena_alloc(cx, sizeof(Com));
    if (cx.status) return 0;
    Com &com = *at<Com>(cx, coff);
    com.zero_id = (uint8_t)h[0].id(); com.nact = com.npend = 0;
    PD_ROLLED
    for (int i = routes.n - 1; i >= 0; --i) {                    // divlist reverses
        const Route *r = at<Route>(cx, routes.r[i]);
        if (r->len == 0) continue;
        const bool active = r->tvoid < 0 || !(h[0].s[r->tvoid] && h[2].s[r->tvoid]);
        if (active) com.act[com.nact++] = routes.r[i]; else com.pend[com.npend++] = routes.r[i];
    }
    // the guards section is reversed too
    com.nguard = (uint8_t)ng;
    PD_ROLLED
    for (int i = 0; i < ng; ++i) { com.gsuit[i] = gs[ng - 1 - i]; com.granks[i] = gr[ng - 1 - i]; }
    return coff;
}

// first-pass (2568-2590) over the active routes with the cards just played
PD_HDN void first_pass(Ctx &cx, const Com &com, const Card *cards_in, int ncards_in, RouteList &routes, Card *skipped, int &nskipped)
{
    TmpFrame tf(cx);
    Card *cards = tmp<Card>(cx, 56), *rest = tmp<Card>(cx, 56);
    int16_t *pos_first = tmp<int16_t>(cx, 53), *pos_third = tmp<int16_t>(cx, 53);
    int *rem_id = tmp<int>(cx, 56), *brk_id = tmp<int>(cx, 56), *ids = tmp<int>(cx, 56);
    uint8_t *is_rem = tmp<uint8_t>(cx, 56), *is_brk = tmp<uint8_t>(cx, 56);
    uint32_t *brk_tricks = tmp<uint32_t>(cx, 56);
    int ncards = ncards_in;
    PD_ROLLED
    for (int i = 0; i < ncards; ++i) cards[i] = cards_in[i];
    bool remaining_nil = ncards == 0;
    routes.n = 0;
    unsigned long long want = 0;
    bool have_want = false;
    PD_ROLLED
    for (int a = 0; a < com.nact && !cx.status; ++a) {
        const uint32_t roff = com.act[a];
        if (remaining_nil) { rl_push_front(cx, routes, roff); ncards = 0; continue; }
        const Route r = *at<Route>(cx, roff);
        // only the cards just played matter: a 53-bit request mask over the card indices
        if (!have_want) {
            want = 0;
            PD_ROLLED
            for (int i = 0; i < ncards; ++i) want |= 1ull << cidx(cards[i]);
            have_want = true;
        }
        if ((route_mask(r) & want) == 0) {
            // none of the cards leads or is third in any trick of this route (94 % of the routes visited): nothing to
            // remove or break, the route stays as it is and the card list is handed on REVERSED (divlist, 2403)
            if (r.len) rl_push_front(cx, routes, roff);
            PD_ROLLED
            for (int i = 0, j = ncards - 1; i < j; ++i, --j) { const Card c = cards[i]; cards[i] = cards[j]; cards[j] = c; }
            continue;
        }
        have_want = false;                                           // cards may leave the list below
        // rem-break-div (2397-2411): first index of each card as a first card / as a third card of the route's tricks
        PD_ROLLED
        for (int i = 0; i < ncards; ++i) { const int ci = cidx(cards[i]); pos_first[ci] = -1; pos_third[ci] = -1; }   // only these are read
        if (r.len > 32767) { fail(cx, PD_E_CAP); return; }
        {
            const uint32_t *tr = at<uint32_t>(cx, r.tricks);
#if defined(__CUDA_ARCH__)
            // 32 tricks per step, one per lane; the (rare) tricks that hold a wanted card are then visited in order by
            // every lane, so all lanes end with the same positions
            PD_ROLLED
            for (uint32_t base = 0; base < r.len; base += 32u) {
                const uint32_t i = base + (uint32_t)PD_LANE;
                const uint32_t t = i < r.len ? tr[i] : 0u;
                const bool hit = i < r.len && (((want >> cidx(tcard(t, 0))) | (want >> cidx(tcard(t, 2)))) & 1ull);
                unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
                PD_ROLLED
                while (bal) {
                    const int b = __ffs((int)bal) - 1;
                    bal &= bal - 1u;
                    const uint32_t tt = __shfl_sync(0xFFFFFFFFu, t, b);
                    const int c0 = cidx(tcard(tt, 0)), c2 = cidx(tcard(tt, 2));
                    if (((want >> c0) & 1ull) && pos_first[c0] < 0) pos_first[c0] = (int16_t)(base + b);
                    if (((want >> c2) & 1ull) && pos_third[c2] < 0) pos_third[c2] = (int16_t)(base + b);
                }
            }
#else
            PD_ROLLED
            for (uint32_t i = 0; i < r.len; ++i) {
                const uint32_t t = tr[i];
                const int c0 = cidx(tcard(t, 0)), c2 = cidx(tcard(t, 2));
                if (((want >> c0) & 1ull) && pos_first[c0] < 0) pos_first[c0] = (int16_t)i;
                if (((want >> c2) & 1ull) && pos_third[c2] < 0) pos_third[c2] = (int16_t)i;
            }
#endif
        }
        const bool no_from = r.frm < 0;
        const bool first_only = no_from || r.frm == r.to;
        // items (card, id) in the three sections, each in REVERSE card order
        PD_ROLLED
        for (int i = 0; i < ncards; ++i) {
            const int ci = cidx(cards[i]);
            const int bid = no_from ? pos_third[ci] : -1;
            int rid;
            if (first_only) rid = cards[i] == NOCARD ? -1 : pos_first[ci];
            else {
                const int p1 = pos_first[ci], p3 = pos_third[ci];
                rid = p1 < 0 ? p3 : (p3 < 0 ? p1 : (p1 < p3 ? p1 : p3));
            }
            is_rem[i] = rid >= 0;
            is_brk[i] = !is_rem[i] && bid >= 0;
            rem_id[i] = bid >= 0 ? bid : rid;          // (or break-id rem-id)
            brk_id[i] = bid;
        }
        // ids to delete; break ids (those of the break section not also in the remove section), in section order
        bool any_ids = false;
        PD_ROLLED
        for (int i = 0; i < ncards; ++i) any_ids = any_ids || is_rem[i] || is_brk[i];
        uint32_t updated = roff;
        uint32_t updated_len = r.len;
        if (any_ids) {
            // del-at (2413-2419): the ids, sorted, then one pass over the tricks
            int nid = 0;
            PD_ROLLED
            for (int i = 0; i < ncards; ++i) {
                if (!is_rem[i] && !is_brk[i]) continue;
                const int id = is_rem[i] ? rem_id[i] : brk_id[i];
                int j = nid++;
                PD_ROLLED
                while (j > 0 && ids[j - 1] > id) { ids[j] = ids[j - 1]; --j; }
                ids[j] = id;
            }
            const uint32_t off = arena_alloc(cx, r.len * 4u);
            if (cx.status) return;
            uint32_t *dst = at<uint32_t>(cx, off);
            const uint32_t *tr = at<uint32_t>(cx, r.tricks);
            uint32_t n = 0;
#if defined(__CUDA_ARCH__)
            // order-preserving compaction, 32 tricks per step: the ids are sorted, `lo` = how many lie below this step
            int lo = 0;
            PD_ROLLED
            for (uint32_t base = 0; base < r.len; base += 32u) {
                const uint32_t t = base + (uint32_t)PD_LANE;
                PD_ROLLED
                while (lo < nid && (uint32_t)ids[lo] < base) ++lo;
                bool keep = t < r.len;
                PD_ROLLED
                for (int k = lo; keep && k < nid && (uint32_t)ids[k] < base + 32u; ++k) if ((uint32_t)ids[k] == t) keep = false;
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) dst[n + __popc(bal & ((1u << PD_LANE) - 1u))] = tr[t];
                n += __popc(bal);
            }
            PD_SYNC();
#else
            int k = 0;
            PD_ROLLED
            for (uint32_t t = 0; t < r.len; ++t) {
                PD_ROLLED
                while (k < nid && (uint32_t)ids[k] < t) ++k;
                if (k < nid && (uint32_t)ids[k] == t) continue;
                dst[n++] = tr[t];
            }
#endif
            updated = new_route(cx, r.frm, r.to, r.tvoid, off, n, route_mask(r));
            updated_len = n;
        }
        int nb = 0;
        PD_ROLLED
        for (int i = ncards - 1; i >= 0; --i) {                       // section order = reverse card order
            if (!is_brk[i]) continue;
            bool in_remove = false;
            PD_ROLLED
            for (int j = 0; j < ncards && !in_remove; ++j) in_remove = is_rem[j] && rem_id[j] == brk_id[i];
            if (!in_remove) brk_tricks[nb++] = at<uint32_t>(cx, r.tricks)[brk_id[i]];
        }
        if (nb) {
            const uint32_t off = tricks_from(cx, brk_tricks, nb);
            insert_route(cx, routes, new_route(cx, r.to, r.to, -1, off, (uint32_t)nb, route_mask(r)));
        }
        if (updated_len) rl_push_front(cx, routes, updated);
        // what is left of the cards: the absent section, reversed
        int nr = 0;
        PD_ROLLED
        for (int i = ncards - 1; i >= 0; --i) if (!is_rem[i] && !is_brk[i]) rest[nr++] = cards[i];
        PD_ROLLED
        for (int i = 0; i < nr; ++i) cards[i] = rest[i];
        ncards = nr;
        if (ncards == 0) remaining_nil = true;
    }
    nskipped = ncards;
    PD_ROLLED
    for (int i = 0; i < ncards; ++i) skipped[i] = cards[i];
}

// after-tricks of a play-com (2643-2660); returns the new play-com
PD_HDN uint32_t com_after_tricks_full(Ctx &cx, uint32_t com_off, const Outcome &out, SearchFrame *stack)
{
    TmpFrame tf(cx);
    Card *raw = tmp<Card>(cx, 56), *skipped = tmp<Card>(cx, 56);
    RouteList &base = *tmp<RouteList>(cx), &new_active = *tmp<RouteList>(cx), &new_pending = *tmp<RouteList>(cx);
    RouteList &act = *tmp<RouteList>(cx), &pend = *tmp<RouteList>(cx);
    int8_t *broken = tmp<int8_t>(cx, MAXG), *ks = tmp<int8_t>(cx, MAXG);
    uint16_t *kr = tmp<uint16_t>(cx, MAXG);
    // normalize-hands (2622-2638)
    PD_PROF_T0();
    uint32_t self_off = com_off;
    {
        const Com &com = *at<Com>(cx, com_off);
        int offset = -1;
        PD_ROLLED
        for (int i = 0; i < 4; ++i) if (out.rem[i].id() == com.zero_id) { offset = i; break; }
        if (offset < 0) { fail(cx, PD_E_LISP); return com_off; }
        if (offset > 0) {
            self_off = arena_alloc(cx, sizeof(Com));
            if (cx.status) return com_off;
            Com &n = *at<Com>(cx, self_off);
            const Com &c = *at<Com>(cx, com_off);
            n.zero_id = (uint8_t)out.rem[0].id(); n.nact = c.nact; n.npend = c.npend; n.nguard = 0;
            PD_ROLLED
            for (int i = 0; i < c.nact; ++i) { const Route r = *at<Route>(cx, c.act[i]); n.act[i] = new_route(cx, r.frm >= 0 ? (r.frm + offset) & 3 : -1, (r.to + offset) & 3, r.tvoid, r.tricks, r.len, route_mask(r)); }
            PD_ROLLED
            for (int i = 0; i < c.npend; ++i) { const Route r = *at<Route>(cx, c.pend[i]); n.pend[i] = new_route(cx, r.frm >= 0 ? (r.frm + offset) & 3 : -1, (r.to + offset) & 3, r.tvoid, r.tricks, r.len, route_mask(r)); }
        }
    }
    if (cx.status) return com_off;
    int nraw = 0;
    PD_ROLLED
    for (int t = 0; t < out.n; ++t) {
        PD_ROLLED
        for (int i = 0; i < 4; ++i) raw[nraw++] = tcard(out.tricks[t], i);
    }
    int nskipped = 0;
    PD_PROF_ADD(5);
    if (at<Com>(cx, self_off)->nact == 0) {                          // no active route: every card is skipped (7 calls in 10)
        base.n = 0; nskipped = nraw;
        PD_ROLLED
        for (int i = 0; i < nraw; ++i) skipped[i] = raw[i];
    } else first_pass(cx, *at<Com>(cx, self_off), raw, nraw, base, skipped, nskipped);
    if (cx.status) return com_off;
    PD_PROF_ADD(6);
    // second-pass (2592-2596): pending routes whose target hand is now void in the route's suit; both lists reversed
    new_active.n = new_pending.n = 0;
    {
        const Com &self = *at<Com>(cx, self_off);
        PD_ROLLED
        for (int i = self.npend - 1; i >= 0; --i) {
            const Route *r = at<Route>(cx, self.pend[i]);
            if (r->tvoid < 0) { fail(cx, PD_E_LISP); return com_off; }
            if (out.rem[r->to].s[r->tvoid] == 0) rl_push(cx, new_active, self.pend[i]); else rl_push(cx, new_pending, self.pend[i]);
        }
    }
    // broken-guards (2598-2615): overwrites the guards of `self` (which is the caller's play-com when offset = 0)
    int nbroken = 0;
    {
        Com &self = *at<Com>(cx, self_off);
        int nk = 0;
        PD_ROLLED
        for (int g = 0; g < self.nguard; ++g) {
            uint16_t played = 0;
            PD_ROLLED
            for (int i = 0; i < nskipped; ++i) if (skipped[i] != NOCARD && csuit(skipped[i]) == self.gsuit[g]) played |= (uint16_t)(1u << crank(skipped[i]));
            const uint16_t left = (uint16_t)(self.granks[g] & ~played);
            if (left) {
                PD_ROLLED
                for (int i = nk; i > 0; --i) { ks[i] = ks[i - 1]; kr[i] = kr[i - 1]; }
                ks[0] = self.gsuit[g]; kr[0] = left; ++nk;
            }
            else {
                PD_ROLLED
                for (int i = nbroken; i > 0; --i) broken[i] = broken[i - 1];
                broken[0] = self.gsuit[g]; ++nbroken;
            }
        }
        self.nguard = (uint8_t)nk;
        PD_ROLLED
        for (int i = 0; i < nk; ++i) { self.gsuit[i] = ks[i]; self.granks[i] = kr[i]; }
    }
    uint32_t refreshed = 0;
    if (nbroken) {
        refreshed = play_routes(cx, out.rem, broken, nbroken, stack);
        if (cx.status) return com_off;
    }
    // the new play-com
    act.n = 0;
    PD_ROLLED
    for (int i = base.n - 1; i >= 0; --i) rl_push(cx, act, base.r[i]);                 // (reverse base-routes)
    PD_ROLLED
    for (int i = 0; i < new_active.n; ++i) insert_route(cx, act, new_active.r[i]);
    pend.n = new_pending.n;
    PD_ROLLED
    for (int i = 0; i < new_pending.n; ++i) pend.r[i] = new_pending.r[i];
    {
        const Com &self = *at<Com>(cx, self_off);                                      // the arena never moves: references stay valid
        if (refreshed) {
            const Com &rf = *at<Com>(cx, refreshed);
            PD_ROLLED
            for (int i = 0; i < rf.nact; ++i) insert_route(cx, act, rf.act[i]);
        }
        PD_ROLLED
        for (int i = 0; i < self.npend; ++i) insert_route(cx, pend, self.pend[i]);
        if (refreshed) {
            const Com &rf = *at<Com>(cx, refreshed);
            PD_ROLLED
            for (int i = 0; i < rf.npend; ++i) insert_route(cx, pend, rf.pend[i]);
        }
        if (cx.status) return com_off;
        const uint32_t noff = arena_alloc(cx, sizeof(Com));
        if (cx.status) return com_off;
        Com &n = *at<Com>(cx, noff);
        n.zero_id = self.zero_id; n.nact = (uint8_t)act.n; n.npend = (uint8_t)pend.n;
        PD_ROLLED
        for (int i = 0; i < act.n; ++i) n.act[i] = act.r[i];
        PD_ROLLED
        for (int i = 0; i < pend.n; ++i) n.pend[i] = pend.r[i];
        int ng = 0;
        PD_ROLLED
        for (int i = 0; i < self.nguard; ++i) { n.gsuit[ng] = self.gsuit[i]; n.granks[ng] = self.granks[i]; ++ng; }
        if (refreshed) {
            const Com &rf = *at<Com>(cx, refreshed);
            PD_ROLLED
            for (int i = 0; i < rf.nguard; ++i) { if (ng >= MAXG) { fail(cx, PD_E_CAP); break; } n.gsuit[ng] = rf.gsuit[i]; n.granks[ng] = rf.granks[i]; ++ng; }
        }
        n.nguard = (uint8_t)ng;
        PD_PROF_ADD(7);
        return noff;
    }
}

// Half of all calls are made on a play-com with no routes and no guards (the strategy has run out of plans): what comes
// back is another empty play-com for the hand now on lead.  Nothing ever writes to such a play-com, so the deal keeps
// four of them (cx.empties, one per hand id) and hands those out: no allocation, and for a play-com that is already one
// of the four not even a load.  That case stays out of the large routine above -- the estimator is bound by instruction
// fetch, and this keeps 30 KB of code out of most nodes' working set.
PD_HDN uint32_t com_after_tricks_loaded(Ctx &cx, uint32_t com_off, const Outcome &out, SearchFrame *stack)
{
    const Com &c = *at<Com>(cx, com_off);
    if ((c.nact | c.npend | c.nguard) != 0 || !cx.empties) return com_after_tricks_full(cx, com_off, out, stack);
    bool found = false;
    PD_ROLLED
    for (int i = 0; i < 4 && !found; ++i) found = out.rem[i].id() == c.zero_id;      // normalize-hands (2622-2638)
    if (!found) { fail(cx, PD_E_LISP); return com_off; }
    return cx.empties + (uint32_t)sizeof(Com) * (uint32_t)out.rem[0].id();
}
PD_HD uint32_t com_after_tricks(Ctx &cx, uint32_t com_off, const Outcome &out, SearchFrame *stack)
{
    if (cx.empties && com_off - cx.empties < 4u * (uint32_t)sizeof(Com))             // one of the four: the hand ids are a
        return cx.empties + (uint32_t)sizeof(Com) * (uint32_t)out.rem[0].id();      // permutation of 0..3, so its hand is there
    return com_after_tricks_loaded(cx, com_off, out, stack);
}
PD_HDN void empties_init(Ctx &cx)
{
    cx.empties = 0;
    const uint32_t off = arena_alloc(cx, 4u * (uint32_t)sizeof(Com));
    if (cx.status) return;
    PD_ROLLED
    for (int k = 0; k < 4; ++k) { Com &n = *at<Com>(cx, off + (uint32_t)sizeof(Com) * k); n.zero_id = (uint8_t)k; n.nact = n.npend = n.nguard = 0; }
    PD_SYNC();
    cx.empties = off;
}

// The search over a deal's play reaches the same four hands by many orders of play: 3 500 nodes for 250 distinct positions
// in an average deal.  A node's value is NOT a function of its hands (it depends on the routes left, the guards and the
// ranks already hunted), but two of the three things tried at every node are: the look-ahead part of play-strategy (its
// immediate-tricks searches and the choice of the target suit: hands, trump and the root strategy's fixed lists) and the
// plain trick in the leader's first suit.  Those are cached per deal, keyed by the four hand words and a tag (root strategy
// 0 / 1, or 2 for the plain trick).  Entries are carved from the END of the arena downwards, so they survive the
// backtracking of the bump allocator at its start; when the two meet, caching stops -- it is never an error.
struct PosEntry { uint64_t key[4]; uint32_t next; uint8_t tag, stage; int8_t target; uint8_t pad; uint32_t pad2[2]; Outcome val; };
constexpr uint32_t POS_BUCKETS = 1024;
PD_HD uint32_t pos_bucket(const Hand *h, int tag)
{
    uint64_t x = h[0].s.m ^ (h[1].s.m << 13 | h[1].s.m >> 51) ^ (h[2].s.m << 29 | h[2].s.m >> 35) ^ (h[3].s.m << 43 | h[3].s.m >> 21);
    x = (x + (uint64_t)tag) * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(x >> 54) & (POS_BUCKETS - 1u);
}
PD_HDN void pos_init(Ctx &cx)
{
    cx.pos = 0;
    if (cx.cap < (1u << 16)) return;                                 // small explicit arenas: no cache
    cx.cap = (cx.cap - POS_BUCKETS * 4u) & ~15u;
    cx.pos = cx.cap;
    uint32_t *b = at<uint32_t>(cx, cx.pos);
    PD_ROLLED
    for (uint32_t i = PD_LANE; i < POS_BUCKETS; i += PD_LANES) b[i] = 0u;
    PD_SYNC();
}
PD_HDN const PosEntry *pos_find(Ctx &cx, const Hand *h, int tag)
{
    if (!cx.pos) return nullptr;
    uint32_t e = at<uint32_t>(cx, cx.pos)[pos_bucket(h, tag)];
    PD_ROLLED
    while (e) {
        const PosEntry *m = at<PosEntry>(cx, e);
        if (m->tag == tag && m->key[0] == h[0].s.m && m->key[1] == h[1].s.m && m->key[2] == h[2].s.m && m->key[3] == h[3].s.m) return m;
        e = m->next;
    }
    return nullptr;
}
PD_HDN void pos_put(Ctx &cx, const Hand *h, int tag, int stage, int target, const Outcome *val)
{
    if (!cx.pos || cx.status) return;
    if (cx.cap < cx.top + (uint32_t)sizeof(PosEntry) + (64u << 10)) return;       // keep 64 KB clear for the routes: stop caching
    cx.cap -= (uint32_t)sizeof(PosEntry);
    const uint32_t off = cx.cap;
    PosEntry *m = at<PosEntry>(cx, off);
    uint32_t *head = at<uint32_t>(cx, cx.pos) + pos_bucket(h, tag);
    PD_ROLLED
    for (int i = 0; i < 4; ++i) m->key[i] = h[i].s.m;
    m->next = *head; m->tag = (uint8_t)tag; m->stage = (uint8_t)stage; m->target = (int8_t)target; m->pad = 0;
    if (val) outcome_copy(m->val, *val);
    PD_SYNC();
    *head = off;
    PD_SYNC();
}

// ---------------------------------------------------------------------------
// play-strategy (2847-2866), hunt-card (2788-2812)
// ---------------------------------------------------------------------------
PD_HDN void hunt_card(Ctx &cx, Outcome &o, int suit, int rank, const Hand *h)
{
    bool top[4] = {false, false, false, false};
    if ((h[1].s[suit] >> rank) & 1u) top[1] = true;
    else if ((h[3].s[suit] >> rank) & 1u) { if (beats(h[1], h[2], suit, cx.trump)) top[0] = true; else top[2] = true; }
    TmpFrame tf(cx);
    Trick &t = *tmp<Trick>(cx);
    trick_init(t, suit, cx.trump);
    PD_ROLLED
    for (int i = 0; i < 4; ++i) beat_or_low(cx, t, h[i], top[i]);
    outcome_of_trick(cx, o, t);
}

// `tag`: which root strategy `rs` is (0 / 1), the position cache's key beside the hands.  Everything up to the hunt
// for a card is a function of the hands, the trump and rs's fixed lists; stage = where that part ended:
// 1 the first search found tricks, 2 no target suit, 3 the second search found tricks, 4 neither (target chosen).
PD_HDN void play_strategy(Ctx &cx, RootStrategy &rs, int tag, const Hand *h, SearchFrame *stack, Outcome &out)
{
    const int trump = cx.trump;
    Target *target = nullptr;
    const PosEntry *known = pos_find(cx, h, tag);
    if (known) {
        if (known->stage == 2) { outcome_empty(out, h); return; }
        if (known->stage != 4) { outcome_copy(out, known->val); return; }
        target = &rs.establish[known->target];
    } else {
        int8_t suits[MAXT + 1];
        int ns = 0;
        PD_ROLLED
        for (int i = 0; i < rs.nsoon; ++i) if (h[0].s[rs.soonest[i]]) suits[ns++] = rs.soonest[i];
        if (ns > 2) ns = 2;                                              // (list-lead 2 suits)
        if (ns) search(cx, IMMEDIATE, -1, suits, ns, h, stack, out);
        else outcome_empty(out, h);                                      // nothing to try: what the search returns for no options
        if (cx.status) return;
        if (out.n > 0) { pos_put(cx, h, tag, 1, -1, &out); return; }
        int ti = -1;
        PD_ROLLED
        for (int i = 0; i < rs.nest && ti < 0; ++i) {
            const int suit = rs.establish[i].suit;
            const bool ok = h[0].s[suit] && (suit != trump || h[1].s[suit] || h[3].s[suit])
                            && (trump < 0 || trump == suit || (h[1].s[suit] && h[3].s[suit]));
            if (ok) ti = i;
        }
        if (ti < 0) { outcome_empty(out, h); pos_put(cx, h, tag, 2, -1, nullptr); return; }
        target = &rs.establish[ti];
        suits[0] = target->suit;
        ns = 1;
        PD_ROLLED
        for (int i = 0; i < rs.nsoon; ++i) suits[ns++] = rs.soonest[i];
        search(cx, IMMEDIATE, -1, suits, ns, h, stack, out);
        if (cx.status) return;
        if (out.n > 0) { pos_put(cx, h, tag, 3, ti, &out); return; }
        pos_put(cx, h, tag, 4, ti, nullptr);
    }
    const int suit = target->suit;
    if (target->head < target->ntheirs) {
        const int rank = target->theirs[target->head++];             // (pop (fourth target))
        hunt_card(cx, out, suit, rank, h);
    } else simple_trick(cx, out, suit, h, false);
}

// ---------------------------------------------------------------------------
// deal-play / play-outcome (2912-2963)
// ---------------------------------------------------------------------------
struct Strategy { uint8_t root; uint32_t com; };
struct PlayFrame {
    Hand hands[4];
    Outcome outcome;             // valid = 0: none yet
    Strategy active, passive;
    uint8_t opt;
    int8_t any_suit;
    Outcome best;
    uint8_t have_best;
    uint32_t mark;
};

struct Result { int status; uint8_t score[2]; uint8_t n; Card tricks[13][4]; uint8_t wno[13]; uint32_t arena_peak; uint32_t nodes; uint32_t tmp_peak; };

// scratch the caller provides (local memory on the device): the stacks of the two kinds of search
struct Scratch {
    SearchFrame sstack[MAXD];
    PlayFrame pstack[MAXD];
    RootStrategy roots[2];
    Result res;
    alignas(16) uint8_t tmp[4736];    // temporaries of the call chain (structural peak 4 320 B: play-deal > after-tricks > play-routes > search > simple-trick)
};

// Probe mode: only the two strategies are set up, and res.nodes receives a score that grows with the expected length of the
// search.  The launcher plays the deals it expects to be long first, because a batch ends with its longest deal and the
// cost of a deal varies by a factor of 100.  The score is a small boosted-tree model over ORDER_FEATURES numbers read off
// the two root strategies and the four hands, fitted on measured device cycles (tools/fit_playdeal_order.py writes
// playdeal_order_model.inc; profiles/r02_playdeal_schedule.md).  It decides the ORDER of play only, never a result.
#if defined(__CUDA_ARCH__)
#define PD_TABLE __device__ const
#else
#define PD_TABLE static const
#endif
constexpr int ORDER_FEATURES = 74;
#include "playdeal_order_model.inc"

PD_HDN void order_features(Ctx &cx, const Scratch &S, uint32_t com_a, uint32_t com_p, const Hand *hands, int *F)
{
    int k = 0;
    PD_ROLLED
    for (int w = 0; w < 2; ++w) {
        const Com &c = *at<Com>(cx, w ? com_p : com_a);
        const RootStrategy &rs = S.roots[w];
        int tl = 0, pl = 0, mx = 0, selfr = 0, nofrom = 0, sd = 0, so = 0, st = 0;
        PD_ROLLED
        for (int i = 0; i < c.nact; ++i) {
            const Route *r = at<Route>(cx, c.act[i]);
            tl += (int)r->len; if ((int)r->len > mx) mx = (int)r->len;
            if (r->frm == r->to) ++selfr;
            if (r->frm < 0) ++nofrom;
        }
        PD_ROLLED
        for (int i = 0; i < c.npend; ++i) pl += (int)at<Route>(cx, c.pend[i])->len;
        PD_ROLLED
        for (int i = 0; i < rs.nest; ++i) { sd += rs.establish[i].diff; so += rs.establish[i].nours; st += rs.establish[i].ntheirs; }
        F[k++] = c.nact; F[k++] = c.npend; F[k++] = c.nguard; F[k++] = tl; F[k++] = pl; F[k++] = mx; F[k++] = selfr; F[k++] = nofrom;
        F[k++] = rs.nsoon; F[k++] = rs.nest; F[k++] = sd; F[k++] = so; F[k++] = st;
    }
    PD_ROLLED
    for (int h = 0; h < 4; ++h) {
        PD_ROLLED
        for (int suit = 0; suit < 4; ++suit) {
            const uint32_t v = hands[h].s[suit];
            F[k++] = popc(v); F[k++] = (int)(v >> 9); F[k++] = (int)(v >> 7);
        }
    }
}
PD_HDN uint32_t order_score(const int *F)
{
    int score = 0;
    PD_ROLLED
    for (int t = 0; t < PD_ORDER_TREES; ++t) {
        int k = PD_ORDER_ROOT[t];
        PD_ROLLED
        while (PD_ORDER_NODE[k][0] >= 0) k = F[PD_ORDER_NODE[k][0]] <= PD_ORDER_NODE[k][1] ? PD_ORDER_NODE[k][2] : PD_ORDER_NODE[k][3];
        score += PD_ORDER_LEAF[k];
    }
    return (uint32_t)(score + (1 << 30));
}

PD_HDN void play_deal(uint8_t *arena, uint32_t arena_bytes, int trump, const Hand *declarer_left_dummy_right, Scratch &S, Result &res,
                      bool probe = false, int *features_out = nullptr)
{
    Ctx cx;
    cx.arena = arena; cx.cap = arena_bytes; cx.top = 64; cx.status = PD_OK; cx.trump = trump;
    cx.sstk = S.tmp; cx.scap = (uint32_t)sizeof(S.tmp); cx.stop = 0; cx.speak = 0; cx.pos = 0; cx.empties = 0;
    res.arena_peak = 0; res.nodes = 0; res.n = 0; res.score[0] = res.score[1] = 0;
    Hand *in = tmp<Hand>(cx, 4), *rolled = tmp<Hand>(cx, 4);
    PD_ROLLED
    for (int i = 0; i < 4; ++i) { in[i] = declarer_left_dummy_right[i]; rolled[i] = declarer_left_dummy_right[(i + 1) & 3]; }   // (roll -1 hands): the left-hand opponent leads
    int8_t *all_suits = tmp<int8_t>(cx, 4);
    PD_ROLLED
    for (int i = 0; i < 4; ++i) all_suits[i] = (int8_t)i;
    // active = strategy of the side on lead, passive = declarer's (2923-2926)
    PD_PROF_T0();
    find_opp_weakness(cx, S.roots[0], rolled, S.sstack);
    const uint32_t com_a = cx.status ? 0 : play_routes(cx, rolled, all_suits, 4, S.sstack);
    if (!cx.status) find_opp_weakness(cx, S.roots[1], in, S.sstack);
    const uint32_t com_p = cx.status ? 0 : play_routes(cx, in, all_suits, 4, S.sstack);
    PD_PROF_ADD(0);
    if (probe) {
        res.status = cx.status; res.tmp_peak = cx.speak; res.arena_peak = cx.top;
        if (cx.status) { res.nodes = 0xFFFFFFFFu; return; }               // outgrew the arena already: a long one
        int *F = tmp<int>(cx, ORDER_FEATURES);
        order_features(cx, S, com_a, com_p, in, F);
        res.nodes = order_score(F);
        if (features_out) { PD_ROLLED for (int i = 0; i < ORDER_FEATURES; ++i) features_out[i] = F[i]; }
        return;
    }
    if (!cx.status) pos_init(cx);
    if (!cx.status) empties_init(cx);
    int sp = 0;
    {
        PlayFrame &f = S.pstack[0];
        PD_ROLLED
        for (int i = 0; i < 4; ++i) f.hands[i] = rolled[i];
        outcome_nil(f.outcome);
        f.active.root = 0; f.active.com = com_a; f.passive.root = 1; f.passive.com = com_p;
        f.opt = 0; f.have_best = 0; outcome_nil(f.best);
    }
    Outcome &final = *tmp<Outcome>(cx), &ans = *tmp<Outcome>(cx);
    outcome_nil(final);
    bool entering = true;
    PD_ROLLED
    while (!cx.status) {
        PlayFrame &f = S.pstack[sp];
        if (entering) {
            ++res.nodes;
            f.any_suit = -1;
            PD_ROLLED
            for (int s = 0; s < 4; ++s) if (f.hands[0].s[s]) { f.any_suit = (int8_t)s; break; }
            f.opt = 0; f.have_best = 0; outcome_nil(f.best);
            entering = false;
            if (f.any_suit < 0) { outcome_copy(f.best, f.outcome); f.opt = 3; }      // (outcome this)
        }
        if (f.opt == 3) {
            if (sp == 0) { outcome_copy(final, f.best); break; }
            // hand the value to the parent: (best-outcome* roll options) (2260-2267)
            PlayFrame &p = S.pstack[sp - 1];
            const int cur = p.outcome.valid ? (p.outcome.roll & 1) : 0;
            if (!p.have_best) { outcome_copy(p.best, f.best); p.have_best = 1; }
            else if (!p.best.valid || (f.best.valid && p.best.score[cur] < f.best.score[cur])) outcome_copy(p.best, f.best);
            if (cx.top > res.arena_peak) res.arena_peak = cx.top;
            cx.top = p.mark;                                         // everything the option allocated is unreachable now
            --sp;
            continue;
        }
        const int option = f.opt++;
        f.mark = cx.top;
        {
            PD_PROF_T0();
            if (option == 0) play_strategy(cx, S.roots[f.active.root], f.active.root, f.hands, S.sstack, ans);
            else if (option == 1) { if (at<Com>(cx, f.active.com)->nact) mk_route(cx, f.active.com, f.hands, ans); else outcome_empty(ans, f.hands); }   // no route to follow: nil
            else {                                                   // the plain trick: a function of the hands
                const PosEntry *known = pos_find(cx, f.hands, 2);
                if (known) outcome_copy(ans, known->val);
                else { simple_trick(cx, ans, f.any_suit, f.hands, false); pos_put(cx, f.hands, 2, 0, -1, &ans); }
            }
            PD_PROF_ADD(1 + option);
        }
        if (cx.status) break;
        if (ans.n == 0) {                                            // after-action gave nil
            if (!f.have_best) { outcome_nil(f.best); f.have_best = 1; }
            cx.top = f.mark;
            continue;
        }
        uint32_t new_a, new_p;
        {
            PD_PROF_T0();
            new_a = com_after_tricks(cx, f.active.com, ans, S.sstack);
            new_p = cx.status ? 0 : com_after_tricks(cx, f.passive.com, ans, S.sstack);
            PD_PROF_ADD(4);
        }
        if (cx.status) break;
        if (sp + 1 >= MAXD) { fail(cx, PD_E_CAP); break; }
        PlayFrame &c = S.pstack[sp + 1];
        PD_ROLLED
        for (int i = 0; i < 4; ++i) c.hands[i] = ans.rem[i];
        if (f.outcome.valid) { outcome_copy(c.outcome, f.outcome); outcome_add(cx, c.outcome, ans, true); } else outcome_copy(c.outcome, ans);
        const bool w = outcome_won(ans);
        Strategy na = {f.active.root, new_a}, np = {f.passive.root, new_p};
        c.active = w ? na : np;
        c.passive = w ? np : na;
        ++sp;
        entering = true;
    }
    if (cx.top > res.arena_peak) res.arena_peak = cx.top;
    res.status = cx.status;
    res.tmp_peak = cx.speak;
    if (!cx.status && !final.valid) res.status = PD_E_NIL;
    if (res.status == PD_OK) {
        res.score[0] = final.score[0]; res.score[1] = final.score[1]; res.n = final.n;
        PD_ROLLED
        for (int t = 0; t < final.n; ++t) {
            PD_ROLLED
            for (int i = 0; i < 4; ++i) res.tricks[t][i] = tcard(final.tricks[t], i);
            res.wno[t] = final.wno[t];
        }
    }
}

}  // namespace pd
