// deal.cuh -- Philox4x32-10 and the in-register 52-card deal (device side).
//
// Replaces the reference's `deal`/`take`/`deal-all` list shuffles
// (bridge.cl:627-658): one thread draws one complete deal as two owner
// bit-planes (4 x 32-bit registers), with no memory traffic and no dynamic
// register indexing.
//
// Algorithm (mirrored bit-for-bit on the CPU by orc_gpu_deal in
// oracle/bmc_oracle.c):
//   * 3 Philox4x32-10 calls (key = seed, counter = (index_lo, index_hi, block, 0))
//     give 12 random words.
//   * Cards are dealt in id order (suit c..s, rank 2..A).  Card p (n = 52-p cards
//     left) goes to seat k with probability remaining_k / n: a bounded draw
//     u in [0,n) is compared with the running prefix sums of the remaining
//     capacities.  This is the sequential form of a Fisher-Yates shuffle of the
//     multiset N^13 E^13 S^13 W^13.
//   * The 51 bounded draws are taken as mixed-radix digits of 12 words: word g
//     yields the digits n = GRP_HI[g] .. GRP_LO[g] by successive 32x32->64
//     multiplies (hi = digit, lo = next fraction).  The final low word is the
//     Lemire remainder of x * N_g, N_g = prod n: if it is below 2^32 mod N_g the
//     draw would be biased and the whole index is VOIDED (p = 0.48 %), so every
//     surviving deal is exactly uniform over the 52!/(13!)^4 deals.
//   * Owner bookkeeping is SWAR on one register Q whose bytes hold
//     128 - prefix_j (j = 0..2): E = Q + u*0x010101 has bit 7 of byte j set iff
//     u >= prefix_j; T7 = ~E & 0x808080 marks the prefixes that shrink.  Five
//     instructions per card: IMAD.WIDE, IMAD, LOP3, 2 x LEA.HI.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

namespace bmc {

// compile-time integer sequence (no <utility>: the same text is compiled by NVRTC)
template <int... I> struct iseq {};
template <int N, int... I> struct make_iseq_impl : make_iseq_impl<N - 1, N - 1, I...> {};
template <int... I> struct make_iseq_impl<0, I...> { using type = iseq<I...>; };
template <int N> using make_iseq = typename make_iseq_impl<N>::type;

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

struct PhiloxKey { uint32_t k0[10], k1[10]; };   // round keys, precomputed on the host

__host__ __device__ inline PhiloxKey philox_key(uint64_t seed)
{
    PhiloxKey k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { k.k0[r] = a; k.k1[r] = b; a += PHILOX_W0; b += PHILOX_W1; }
    return k;
}

__device__ __forceinline__ void philox4x32_10(const PhiloxKey &key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// digit groups (n = cards left when the card is dealt)
__host__ __device__ constexpr int grp_hi(int g)
{
    constexpr int t[12] = {52, 48, 44, 40, 36, 32, 28, 24, 20, 16, 11, 5};
    return t[g];
}
__host__ __device__ constexpr int grp_lo(int g)
{
    constexpr int t[12] = {49, 45, 41, 37, 33, 29, 25, 21, 17, 12, 6, 2};
    return t[g];
}
__host__ __device__ constexpr int grp_of(int n)
{
    for (int g = 0; g < 12; ++g) if (n <= grp_hi(g) && n >= grp_lo(g)) return g;
    return -1;
}
__host__ __device__ constexpr uint32_t grp_thresh(int g)
{
    uint64_t N = 1;
    for (int n = grp_lo(g); n <= grp_hi(g); ++n) N *= (uint64_t)n;
    return (uint32_t)((1ull << 32) % N);
}

// 32x32 -> 64 multiply split into (lo, hi).  Written in PTX so that the compiler
// keeps one IMAD.WIDE per digit (the C form picks up a redundant mask and add).
__device__ __forceinline__ void mul_wide(uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi)
{
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}

struct Planes { uint32_t lo0, lo1, hi0, hi1; };   // lo plane (c,d | h,s), hi plane (c,d | h,s)

struct DealState {
    uint32_t w[4];      // current Philox block
    uint32_t x;         // running fraction of the current word
    uint32_t Q;         // bytes 128 - prefix_j
    uint32_t acc;       // t-bit history of the current byte group
    uint32_t lo[2], hi[2];
    bool voided;
};

// One card.  P is a true compile-time constant so that every table lookup, the
// Lemire threshold and all shift counts fold to immediates.
template <int P>
__device__ __forceinline__ void deal_step(const PhiloxKey &key, uint32_t idx_lo, uint32_t idx_hi, DealState &s)
{
    constexpr int n = 52 - P;
    uint32_t u = 0;
    if constexpr (n >= 2) {
        constexpr int g = grp_of(n);
        if constexpr (n == grp_hi(g)) {
            if constexpr ((g & 3) == 0) philox4x32_10(key, idx_lo, idx_hi, (uint32_t)(g >> 2), 0u, s.w);
            s.x = s.w[g & 3];
        }
        mul_wide(s.x, (uint32_t)n, s.x, u);
        if constexpr (n == grp_lo(g)) {
            constexpr uint32_t thresh = grp_thresh(g);
            s.voided |= s.x < thresh;
        }
    }
    const uint32_t E = u * 0x010101u + s.Q;
    const uint32_t T7 = ~E & 0x808080u;          // bit 7 of byte j: u < prefix_j
    s.Q += T7 >> 7;
    constexpr int rank = P % 13, suit = P / 13;
    constexpr int i = rank < 8 ? rank : rank - 8;   // position inside the byte group
    s.acc += T7 >> (7 - i);
    if constexpr (rank == 7 || rank == 12) {
        constexpr uint32_t mask = rank == 7 ? 0xFFu : 0x1Fu;
        const uint32_t par = s.acc ^ (s.acc >> 8) ^ (s.acc >> 16);
        const uint32_t lob = ~par & mask;               // owner bit 0 = !(t0^t1^t2)
        const uint32_t hib = ~(s.acc >> 8) & mask;      // owner bit 1 = !t1
        constexpr int sh = 16 * (suit & 1) + (rank == 7 ? 0 : 8);
        s.lo[suit >> 1] |= lob << sh;
        s.hi[suit >> 1] |= hib << sh;
        s.acc = 0;
    }
}

template <int... P>
__device__ __forceinline__ void deal_all_steps(const PhiloxKey &key, uint32_t idx_lo, uint32_t idx_hi, DealState &s,
                                               iseq<P...>)
{
    (deal_step<P>(key, idx_lo, idx_hi, s), ...);
}

// Returns true when the index is void.
__device__ __forceinline__ bool deal52(const PhiloxKey &key, uint32_t idx_lo, uint32_t idx_hi, Planes &pl)
{
    DealState s;
    s.x = 0;
    s.Q = (128u - 13u) | ((128u - 26u) << 8) | ((128u - 39u) << 16);
    s.acc = 0;
    s.lo[0] = s.lo[1] = s.hi[0] = s.hi[1] = 0u;
    s.voided = false;
    deal_all_steps(key, idx_lo, idx_hi, s, make_iseq<52>{});
    pl.lo0 = s.lo[0]; pl.lo1 = s.lo[1]; pl.hi0 = s.hi[0]; pl.hi1 = s.hi[1];
    return s.voided;
}

// ---------------------------------------------------------------------------
// Deals with fixed (`:=`) hands: the 52 - 13k free cards are dealt to the free
// seats with the same owner bookkeeping, in a run-time loop (the free-card
// positions are per-constraint constants, uniform across the grid).  Word w of
// the Philox stream yields the digits n = n_free - 4w .. n_free - 4w - 3.
// Mirrored by orc_gpu_deal_fixed in oracle/bmc_oracle.c.
// ---------------------------------------------------------------------------
struct FixDev {
    uint32_t n_free;              // 52 = no fixed hand (deal52 is used instead)
    uint32_t q0;                  // initial Q: bytes 128 - prefix_j of the seat capacities
    uint32_t lo0, lo1, hi0, hi1;  // owner planes of the fixed cards
    uint32_t free0, free1;        // lane-layout masks of the free cards
    uint32_t thresh[10];          // Lemire threshold of every 4-digit word
};

#ifndef __CUDACC_RTC__
__host__ inline void fixdev_init(FixDev &f, const uint64_t fixed_mask[4], const bool is_fixed[4])
{
    uint64_t lo = 0, hi = 0, used = 0;
    uint32_t cap[4];
    for (int k = 0; k < 4; ++k) {
        cap[k] = is_fixed[k] ? 0u : 13u;
        if (!is_fixed[k]) continue;
        used |= fixed_mask[k];
        if (k & 1) lo |= fixed_mask[k];
        if (k & 2) hi |= fixed_mask[k];
    }
    const uint64_t free_mask = 0x1FFF1FFF1FFF1FFFull & ~used;
    f.n_free = (uint32_t)__builtin_popcountll(free_mask);
    const uint32_t p0 = cap[0], p1 = p0 + cap[1], p2 = p1 + cap[2];
    f.q0 = (128u - p0) | ((128u - p1) << 8) | ((128u - p2) << 16);
    f.lo0 = (uint32_t)lo; f.lo1 = (uint32_t)(lo >> 32); f.hi0 = (uint32_t)hi; f.hi1 = (uint32_t)(hi >> 32);
    f.free0 = (uint32_t)free_mask; f.free1 = (uint32_t)(free_mask >> 32);
    for (int w = 0; w < 10; ++w) {
        uint64_t N = 1;
        for (int d = 0; d < 4; ++d) {
            const int n = (int)f.n_free - 4 * w - d;
            if (n >= 2) N *= (uint64_t)n;
        }
        f.thresh[w] = (uint32_t)((1ull << 32) % N);
    }
}
#endif

// Core loop.  All arguments but (idx, masks of the v2 path) are grid-uniform; the
// seat-first sampler calls it with PER-LANE masks (the primary seat's hand).
__device__ __forceinline__ bool deal_fixed_core(const PhiloxKey &key, uint32_t stream, uint32_t n_free, uint32_t q0,
                                                uint32_t lo0, uint32_t lo1, uint32_t hi0, uint32_t hi1, uint32_t free0,
                                                uint32_t free1, const uint32_t *__restrict__ thresh, uint32_t idx_lo,
                                                uint32_t idx_hi, Planes &pl)
{
    uint32_t w[4];
    uint32_t x = 0, Q = q0;
    bool voided = false;
    uint32_t n = n_free;
    for (uint32_t v = 0; n >= 1; ++v, --n) {
        uint32_t pos;
        if (free0) { pos = __ffs(free0) - 1; free0 &= free0 - 1; }
        else { pos = 32 + __ffs(free1) - 1; free1 &= free1 - 1; }
        uint32_t u = 0;
        if (n >= 2) {
            if ((v & 3u) == 0u) {
                if ((v & 15u) == 0u) philox4x32_10(key, idx_lo, idx_hi, v >> 4, stream, w);
                const uint32_t sel = (v >> 2) & 3u;
                x = sel == 0u ? w[0] : sel == 1u ? w[1] : sel == 2u ? w[2] : w[3];
            }
            mul_wide(x, n, x, u);
            if ((v & 3u) == 3u || n == 2u) voided |= x < thresh[v >> 2];
        }
        const uint32_t E = u * 0x010101u + Q;
        const uint32_t A = E & 0x808080u;                  // bit 7 of byte j: u >= prefix_j
        Q += (A ^ 0x808080u) >> 7;
        const uint32_t hib = (A >> 15) & 1u;
        const uint32_t lob = ((A >> 7) ^ (A >> 15) ^ (A >> 23)) & 1u;
        if (pos < 32u) { lo0 |= lob << pos; hi0 |= hib << pos; }
        else { lo1 |= lob << (pos - 32u); hi1 |= hib << (pos - 32u); }
    }
    pl.lo0 = lo0; pl.lo1 = lo1; pl.hi0 = hi0; pl.hi1 = hi1;
    return voided;
}

__device__ __forceinline__ bool deal_fixed(const PhiloxKey &key, const FixDev &f, uint32_t idx_lo, uint32_t idx_hi, Planes &pl)
{
    return deal_fixed_core(key, 0u, f.n_free, f.q0, f.lo0, f.lo1, f.hi0, f.hi1, f.free0, f.free1, f.thresh, idx_lo, idx_hi, pl);
}

}  // namespace bmc
