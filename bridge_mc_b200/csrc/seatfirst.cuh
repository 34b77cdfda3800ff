// seatfirst.cuh -- lazy, seat-first dealing for rejection sampling (device).
//
// A uniformly random deal factorises as
//     P(primary seat's honour counts) x P(its cards | counts) x P(other three hands | primary hand)
// and a constraint that rejects most deals on the primary seat alone never needs the later
// factors.  The sampler therefore runs in stages with a warp-level compaction queue between them, so
// every stage executes on full warps:
//   A1  the RANK of the primary seat's hand in an order that puts the honour holdings with an
//       admissible HCP first (a quarter of a Philox call, one 32-bit compare; the low word of the
//       rank is drawn only for ties, 2^-32 per index)                          -> HCP range
//   A2  unrank the hand from the same uniform: honour counts, their suits, the low cards suit by
//       suit (no further random numbers)                                       -> all ranges of the seat
//   B   the other 39 cards to the other three seats (3 Philox calls, digits + Lemire rejection as in
//       deal.cuh)                                                              -> every seat's full test
// Every factor is sampled exactly, so accepted deals are exactly uniform over the deals that satisfy
// the constraint -- the same distribution as dealing all 52 cards and rejecting.  Mirrored bit for
// bit by orc_gpu_seatfirst_hand / gpu_deal_fixed_stream in oracle/bmc_oracle.c.
//
// Stage A in detail.  A 13-card hand holds nJ jacks, nQ queens, nK kings and nA aces with
// probability w / C(52,13), w = I C(36, m), I = C(4,nJ) C(4,nQ) C(4,nK) C(4,nA), m = 13 - nJ - nQ - nK -
// nA, and its HCP is nJ + 2 nQ + 3 nK + 4 nA: the 625 count classes are laid out on [0, C(52,13)) with
// the classes whose HCP lies in the seat's range FIRST.  One 64-bit uniform u then decides the HCP
// test by a single compare (u < theta = S * weight of the admitted classes, S = floor(2^64 /
// C(52,13))); u >= S C(52,13) (probability 1.2e-8) voids the index, so every hand keeps its exact
// probability.  Only the survivors look their class up; q = (u - S cum) / S is uniform on [0, I C(36,m)):
// q mod I picks the suits of the honours, q / I is the rank of the m low cards among the C(36,m)
// possibilities, unranked suit by suit with binomial prefix sums.
#pragma once
#include "deal.cuh"

namespace bmc {

constexpr uint32_t STREAM_SEATFIRST_B = 3u;       // (stream 2 belonged to the card-by-card stage A of earlier builds)

// ---- stage A1, rank form -----------------------------------------------------------
constexpr uint32_t STREAM_SEATFIRST_RANK = 4u, STREAM_SEATFIRST_RANK_LO = 5u;   // hi and lo word of u
constexpr int RANK_TAB_ENTRIES = 640;       // 625 classes at most, then entries no u reaches
constexpr int SF_BUCKETS = 256;

constexpr unsigned long long SF_HANDS = 635013559600ull;          // C(52,13)
constexpr unsigned long long SF_S = 29049370ull;                  // floor(2^64 / C(52,13))
constexpr unsigned long long SF_S_RECIP = 635013567375ull;        // floor(2^64 / SF_S)
static_assert(SF_S * SF_HANDS == 0xFFFFFFCB68F946E0ull, "S C(52,13) is the void threshold");
static_assert(SF_HANDS * (SF_S + 1ull) < SF_HANDS * SF_S, "S = floor(2^64 / C(52,13)): the next multiple wraps");

// Low-card tables (compile-time constants): a hand holds k of the 36 low cards (2..10, nine per suit)
// as l_c, l_d, l_h, l_s of each suit with weight C(9,l_c) C(9,l_d) C(9,l_h) C(9,l_s).
//   prefix[st][k][l] = sum_{j<l} C(9,j) C(R,k-j), R = 27, 18, 9 for st = 0, 1, 2: the number of ways with
//                      fewer than l cards in suit st when k low cards go to suits st..3 (0xFFFFFFFF past the end);
//   c9[l] = { C(9,l), floor(2^32 / C(9,l)), index of the first l-subset in subset9, 0 };
//   subset9 = the 512 subsets of nine cards ordered by size, then by value.
struct alignas(16) LowTabs {
    uint32_t prefix[3][14][16];
    uint32_t c9[10][4];
    uint16_t subset9[512];
};
__host__ __device__ constexpr unsigned long long sf_binom(int n, int k)
{
    if (k < 0 || k > n) return 0ull;
    unsigned long long r = 1ull;
    for (int i = 1; i <= k; ++i) r = r * (unsigned long long)(n - k + i) / (unsigned long long)i;
    return r;
}
__host__ __device__ constexpr LowTabs make_lowtabs()
{
    LowTabs t{};
    for (int st = 0; st < 3; ++st)
        for (int k = 0; k < 14; ++k) {
            const int R = 27 - 9 * st;
            // the running total reaches C(R+9,k) > ell at l = 10 at the latest; entries after it are 0xFFFFFFFF
            unsigned long long run = 0ull;
            for (int l = 0; l < 16; ++l) {
                t.prefix[st][k][l] = l <= 10 ? (uint32_t)(run > 0xFFFFFFFFull ? 0xFFFFFFFFull : run) : 0xFFFFFFFFu;
                if (l <= 9) run += sf_binom(9, l) * sf_binom(R, k - l);
            }
        }
    int pos = 0;
    for (int l = 0; l <= 9; ++l) {
        const unsigned long long c = sf_binom(9, l);
        t.c9[l][0] = (uint32_t)c;
        t.c9[l][1] = c == 1ull ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / c);
        t.c9[l][2] = (uint32_t)pos;
        t.c9[l][3] = 0u;
        for (int v = 0; v < 512; ++v) {
            int pc = 0;
            for (int b = 0; b < 9; ++b) pc += (v >> b) & 1;
            if (pc == l) t.subset9[pos++] = (uint16_t)v;
        }
    }
    return t;
}

// High halves of products; host versions so that the unranking arithmetic can be checked exhaustively on
// the CPU (tests/unrank_selftest.cu).
__host__ __device__ __forceinline__ uint32_t sf_mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}
__host__ __device__ __forceinline__ unsigned long long sf_mulhi64(unsigned long long a, unsigned long long b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
#endif
}

// Table entry (uint4): x,y = S * (weight of the classes before this one) as lo,hi;
// z = nJ | nQ<<3 | nK<<6 | nA<<9 | I<<12;  w = floor(2^32 / I) (2^32 - 1 for I = 1).
// The table is followed by SF_BUCKETS 16-bit entries: bucket[j] = the class holding the smallest u with
// umulhi(u_hi, bmul) = j (bmul ~ SF_BUCKETS 2^32 / (theta_hi + 1)), so a lookup is one indexed load and a
// short forward scan instead of a binary search.
// Finds the class of u; returns its info word, sets d = u - S cum (uniform on [0, S I C(36,m))).
__device__ __forceinline__ uint32_t sf_find_class(const uint4 *__restrict__ tab, uint32_t bmul, uint32_t u_lo, uint32_t u_hi,
                                                  unsigned long long &d, uint32_t &magic)
{
    const unsigned long long u = ((unsigned long long)u_hi << 32) | u_lo;
    const uint16_t *bucket = reinterpret_cast<const uint16_t *>(tab + RANK_TAB_ENTRIES);
    uint32_t pos = __ldg(bucket + min(sf_mulhi32(u_hi, bmul), (uint32_t)SF_BUCKETS - 1u));
    for (;;) {
        const uint2 c = __ldg(reinterpret_cast<const uint2 *>(tab + pos + 1u));
        if ((((unsigned long long)c.y << 32) | c.x) > u || pos >= (uint32_t)RANK_TAB_ENTRIES - 2u) break;
        ++pos;
    }
    const uint4 e = __ldg(tab + pos);
    d = u - ((((unsigned long long)e.y) << 32) | e.x);
    magic = e.w;
    return e.z;
}

// d uniform on [0, S I L): q = d / S is uniform on [0, I L); i = q mod I picks the identity of the
// honours, ell = q / I the low cards -- exactly uniform and independent.  q < 2^42, I <= 1296.
__host__ __device__ __forceinline__ void sf_split(unsigned long long d, uint32_t I, uint32_t magic, uint32_t &i, uint32_t &ell)
{
    unsigned long long q = sf_mulhi64(d, SF_S_RECIP);     // floor(d / S) or one less (tests/unrank_selftest.cu)
    unsigned long long r = d - q * SF_S;
    if (r >= SF_S) { r -= SF_S; ++q; }
    if (r >= SF_S) { r -= SF_S; ++q; }
    const uint32_t qh = (uint32_t)(q >> 21), ql = (uint32_t)q & 0x1FFFFFu;     // long division in base 2^21
    uint32_t a = sf_mulhi32(qh, magic), ra = qh - a * I;
    if (ra >= I) { ra -= I; ++a; }
    const uint32_t t = (ra << 21) | ql;                   // < 1296 * 2^21 < 2^32
    uint32_t b = sf_mulhi32(t, magic);
    i = t - b * I;
    if (i >= I) { i -= I; ++b; }
    ell = (a << 21) + b;
}

// Identity of the honours: i in [0, I) as mixed-radix digits (J first) over the C(4, n.) subsets of suits.
// Returns H: bit 4*suit + level (J Q K A = 0..3) set where the seat holds that honour.
__host__ __device__ __forceinline__ uint32_t sf_honours_of(uint32_t info, uint32_t i)
{
    uint32_t H = 0;
#pragma unroll
    for (int lev = 0; lev < 4; ++lev) {
        const uint32_t c = (info >> (3 * lev)) & 7u;                  // honours of this level in the hand
        const uint32_t r = (0x14641u >> (4u * c)) & 15u;              // C(4, c)
        const uint32_t M = r == 1u ? 65536u : (r == 4u ? 16384u : 10923u);
        const uint32_t q = (i * M) >> 16;                             // i / r, exact for i < 1296
        const uint32_t dg = i - q * r;
        i = q;
        const uint32_t off = (0xFB510u >> (4u * c)) & 15u;            // first subset of size c in the nibble list
        const uint32_t m4 = (uint32_t)(0xFEDB7CA965384210ull >> (4u * (off + dg))) & 15u;   // which suits
        const uint32_t sp = (((m4 * 0x41u) & 0x303u) * 9u) & 0x1111u; // spread: suit s -> bit 4 s (no carries)
        H |= sp << lev;
    }
    return H;
}

// The k low cards of the hand from their rank ell in [0, C(36,k)): suit by suit (c d h s), the length
// by a search over prefix[st][k], the cards by the subset table.  Returns the two lane-layout words
// (low nine bits of each 16-bit lane).
__host__ __device__ __forceinline__ void sf_unrank_lows(const LowTabs &T, uint32_t ell, uint32_t k, uint32_t &m0, uint32_t &m1)
{
    uint32_t sub[4];
    k = k < 13u ? k : 13u;
#pragma unroll
    for (int st = 0; st < 3; ++st) {
        const uint32_t *row = T.prefix[st][k];
        uint32_t l = 0;
#pragma unroll
        for (int s = 8; s != 0; s >>= 1) if (row[l + s] <= ell) l += s;
        l = l < 9u ? l : 9u;
        ell -= row[l];
        const uint4 e = *reinterpret_cast<const uint4 *>(T.c9[l]);
        uint32_t q = sf_mulhi32(ell, e.y), r = ell - q * e.x;          // ell = q C(9,l) + r
        if (r >= e.x) { r -= e.x; ++q; }
        sub[st] = T.subset9[(e.z + r) & 511u];
        ell = q;
        k = k - l < 13u ? k - l : 13u;
    }
    const uint32_t l3 = k < 9u ? k : 9u;
    sub[3] = T.subset9[(T.c9[l3][2] + ell) & 511u];
    m0 = sub[0] | (sub[1] << 16);
    m1 = sub[2] | (sub[3] << 16);
}

// The primary seat's hand behind index `idx` (stage A2; also bmc_expand): u from the two Philox calls, its
// count class, the identity of the honours, the low cards.  Lanes with active = false look up u = 0
// (class 0, nothing to scan) and return that hand.
__device__ __forceinline__ void sf_primary_hand(const PhiloxKey &key, const uint4 *__restrict__ tab, uint32_t bmul, const LowTabs &T,
                                                unsigned long long idx, bool active, uint32_t &m0, uint32_t &m1)
{
    uint32_t wh[4], wl[4];
    philox4x32_10(key, (uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, STREAM_SEATFIRST_RANK, wh);
    philox4x32_10(key, (uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, STREAM_SEATFIRST_RANK_LO, wl);
    const uint32_t j = (uint32_t)idx & 3u;
    const uint32_t u_hi = j < 2u ? (j == 0u ? wh[0] : wh[1]) : (j == 2u ? wh[2] : wh[3]);
    const uint32_t u_lo = j < 2u ? (j == 0u ? wl[0] : wl[1]) : (j == 2u ? wl[2] : wl[3]);
    unsigned long long d;
    uint32_t magic, i, ell;
    const uint32_t info = sf_find_class(tab, bmul, active ? u_lo : 0u, active ? u_hi : 0u, d, magic);
    sf_split(d, (info >> 12) & 0x7FFu, magic, i, ell);
    const uint32_t H = sf_honours_of(info, i);
    sf_unrank_lows(T, ell, 13u - (uint32_t)__popc(H), m0, m1);
    m0 |= ((H & 0xFu) << 9) | ((H & 0xF0u) << 21);
    m1 |= ((H & 0xF00u) << 1) | ((H & 0xF000u) << 13);
}

// ---- stage A, exact form (mode 4) --------------------------------------------------------------------
// The class table built by distgen.cuh lists every class of hands the primary seat's constraint admits --
// (honours exactly, low-card counts per suit) -- with the number of hands before it.  The admitted hands
// occupy the ranks [0, W) of the 64-bit uniform's scale, so stage A1's one compare (u < W S) accepts exactly
// the indices whose primary hand satisfies the seat, and stage A2 only unranks: rank R = u / S, class by a
// bucket index + short binary search on the cumulative counts, then r = R - cum is split into one subset
// index per suit (mixed radix over C(9, l_s)).  Every hand keeps its S preimages in u: exactly uniform.
struct ExactTab {
    const unsigned long long *cum;    // [n_cls + 1] hands before class i
    const uint32_t *keys;             // [n_cls] honours (bit 4 s + level) | l_c << 16 | l_d << 20 | l_h << 24 | l_s << 28
    const uint32_t *bkt;              // [n_bkt + 1] class of the smallest u of bucket j = umulhi(u_hi, bmul)
    uint32_t bmul, n_bkt;
};

__device__ __forceinline__ void sf_exact_hand(const PhiloxKey &key, const ExactTab &x, const LowTabs &T, unsigned long long idx,
                                              bool active, uint32_t &m0, uint32_t &m1)
{
    uint32_t wh[4], wl[4];
    philox4x32_10(key, (uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, STREAM_SEATFIRST_RANK, wh);
    philox4x32_10(key, (uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, STREAM_SEATFIRST_RANK_LO, wl);
    const uint32_t j = (uint32_t)idx & 3u;
    uint32_t u_hi = j < 2u ? (j == 0u ? wh[0] : wh[1]) : (j == 2u ? wh[2] : wh[3]);
    uint32_t u_lo = j < 2u ? (j == 0u ? wl[0] : wl[1]) : (j == 2u ? wl[2] : wl[3]);
    if (!active) { u_hi = 0u; u_lo = 0u; }
    const unsigned long long u = ((unsigned long long)u_hi << 32) | u_lo;
    unsigned long long R = sf_mulhi64(u, SF_S_RECIP);          // floor(u / S) or one less (tests/unrank_selftest.cu)
    unsigned long long rem = u - R * SF_S;
    if (rem >= SF_S) { rem -= SF_S; ++R; }
    if (rem >= SF_S) { rem -= SF_S; ++R; }
    const uint32_t b = min(sf_mulhi32(u_hi, x.bmul), x.n_bkt - 1u);
    uint32_t lo = __ldg(x.bkt + b), hi = __ldg(x.bkt + b + 1u);
    while (lo < hi) {                                           // largest class with cum <= R
        const uint32_t mid = (lo + hi + 1u) >> 1;
        if (__ldg(x.cum + mid) <= R) lo = mid; else hi = mid - 1u;
    }
    const uint32_t k = __ldg(x.keys + lo);
    uint32_t r = (uint32_t)(R - __ldg(x.cum + lo));             // < prod C(9, l_s) <= 126^4 < 2^32
    uint32_t sub[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t l = min((k >> (16 + 4 * s)) & 15u, 9u);
        const uint4 e = *reinterpret_cast<const uint4 *>(T.c9[l]);
        uint32_t q = sf_mulhi32(r, e.y), t = r - q * e.x;       // r = q C(9,l) + t
        if (t >= e.x) { t -= e.x; ++q; }
        sub[s] = T.subset9[(e.z + t) & 511u];
        r = q;
    }
    m0 = sub[0] | (sub[1] << 16) | ((k & 0xFu) << 9) | ((k & 0xF0u) << 21);
    m1 = sub[2] | (sub[3] << 16) | ((k & 0xF00u) << 1) | ((k & 0xF000u) << 13);
}

// Lemire thresholds of deal_fixed's words for 39 free cards (4 digits per word)
__host__ __device__ constexpr uint32_t thresh39(int w)
{
    uint64_t N = 1;
    for (int d = 0; d < 4; ++d) {
        const int n = 39 - 4 * w - d;
        if (n >= 2) N *= (uint64_t)n;
    }
    return (uint32_t)((1ull << 32) % N);
}

// ---------------------------------------------------------------------------
// Stage B, fast form: the other 39 cards are dealt to the three other seats on COMPACT positions
// 0..38 with the static SWAR bookkeeping of deal52 (no per-lane indexing at all), and the two
// compact owner planes are then deposited into the free positions of the primary seat's mask
// (a software pdep, sheep-and-goats expand).  Same digits, same order, same owner rule as
// deal_fixed_core with the primary hand fixed -- hence the same CPU mirror -- at a third of the
// instructions.
// ---------------------------------------------------------------------------
struct RestState {
    uint32_t w[4];
    uint32_t x, Q, acc;
    uint32_t lo[2], hi[2];      // compact planes: word 0 = cards 0..31, word 1 = cards 32..38
    bool voided;
};

template <int V>
__device__ __forceinline__ void rest_step(const PhiloxKey &key, uint32_t idx_lo, uint32_t idx_hi, RestState &s)
{
    constexpr int n = 39 - V;
    uint32_t u = 0;
    if constexpr (n >= 2) {
        if constexpr ((V & 3) == 0) {
            if constexpr ((V & 15) == 0) philox4x32_10(key, idx_lo, idx_hi, (uint32_t)(V >> 4), STREAM_SEATFIRST_B, s.w);
            s.x = s.w[(V >> 2) & 3];
        }
        mul_wide(s.x, (uint32_t)n, s.x, u);
        if constexpr ((V & 3) == 3 || n == 2) {
            constexpr uint32_t thresh = thresh39(V >> 2);
            s.voided |= s.x < thresh;
        }
    }
    const uint32_t E = u * 0x010101u + s.Q;
    const uint32_t T7 = ~E & 0x808080u;
    s.Q += T7 >> 7;
    constexpr int i = V & 7;
    s.acc += T7 >> (7 - i);
    if constexpr (i == 7 || V == 38) {
        constexpr uint32_t mask = V == 38 ? 0x7Fu : 0xFFu;
        const uint32_t par = s.acc ^ (s.acc >> 8) ^ (s.acc >> 16);
        const uint32_t lob = ~par & mask;
        const uint32_t hib = ~(s.acc >> 8) & mask;
        constexpr int word = V >> 5, sh = (V & 31) & ~7;
        s.lo[word] |= lob << sh;
        s.hi[word] |= hib << sh;
        s.acc = 0;
    }
}

template <int... V>
__device__ __forceinline__ void rest_steps(const PhiloxKey &key, uint32_t idx_lo, uint32_t idx_hi, RestState &s,
                                           iseq<V...>)
{
    (rest_step<V>(key, idx_lo, idx_hi, s), ...);
}

// sheep-and-goats masks of a 32-bit deposit (Hacker's Delight 7-5): computed once per mask word,
// applied to both planes
struct Deposit32 { uint32_t mv[5]; uint32_t m; };

__device__ __forceinline__ Deposit32 deposit_prepare(uint32_t m)
{
    Deposit32 d;
    d.m = m;
    uint32_t mk = ~m << 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        uint32_t mp = mk ^ (mk << 1);
        mp ^= mp << 2;
        mp ^= mp << 4;
        mp ^= mp << 8;
        mp ^= mp << 16;
        const uint32_t mv = mp & m;
        d.mv[i] = mv;
        m = (m ^ mv) | (mv >> (1 << i));
        mk &= ~mp;
    }
    return d;
}

__device__ __forceinline__ uint32_t deposit_apply(const Deposit32 &d, uint32_t x)
{
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        const uint32_t t = x << (1 << i);
        x = (x & ~d.mv[i]) | (t & d.mv[i]);
    }
    return x & d.m;
}

// m0, m1: the primary seat's hand (lane layout); q0: initial Q with capacity 0 for that seat;
// (xl, xh): all-ones where the primary seat's owner bit 0 / bit 1 is set.
__device__ __forceinline__ bool deal_rest39(const PhiloxKey &key, uint32_t q0, uint32_t m0, uint32_t m1, uint32_t xl, uint32_t xh,
                                            uint32_t idx_lo, uint32_t idx_hi, Planes &pl)
{
    RestState s;
    s.x = 0; s.Q = q0; s.acc = 0; s.voided = false;
    s.lo[0] = s.lo[1] = s.hi[0] = s.hi[1] = 0u;
    rest_steps(key, idx_lo, idx_hi, s, make_iseq<39>{});
    // split the 39 compact bits between the two words of the lane layout
    const uint32_t f0 = ~m0 & 0x1FFF1FFFu, f1 = ~m1 & 0x1FFF1FFFu;
    const uint32_t n0 = __popc(f0);                                   // 13..26 free cards in clubs + diamonds
    const uint32_t keep = (1u << n0) - 1u;                            // n0 <= 26
    const uint32_t lo_a = s.lo[0] & keep, hi_a = s.hi[0] & keep;
    const uint32_t lo_b = __funnelshift_r(s.lo[0], s.lo[1], n0), hi_b = __funnelshift_r(s.hi[0], s.hi[1], n0);
    const Deposit32 d0 = deposit_prepare(f0), d1 = deposit_prepare(f1);
    pl.lo0 = deposit_apply(d0, lo_a) | (m0 & xl);
    pl.hi0 = deposit_apply(d0, hi_a) | (m0 & xh);
    pl.lo1 = deposit_apply(d1, lo_b) | (m1 & xl);
    pl.hi1 = deposit_apply(d1, hi_b) | (m1 & xh);
    return s.voided;
}

}  // namespace bmc
