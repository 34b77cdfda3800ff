// reforder.cuh -- "reference order" sampler: the reference's own multi-seat semantics.
//
// `build` (bridge.cl:1292-1319) does NOT draw uniformly over the deals that satisfy every seat's
// ranges.  It samples the ranged seats one after the other: seat j is uniform over the 13-card
// hands of the REMAINING cards that satisfy ranges narrowed by infer-distgen-params
// (bridge.cl:1126-1269) against the remaining supply and the later defs, with these quirks:
//   * a suit whose narrowed spec has no explicit upper bound is capped at the number of LOW
//     cards (2..10) left in that suit (bridge.cl:1000-1001 `(or max supply)`);
//   * the last ranged seat is never sampled when only 13 cards remain (bridge.cl:1303);
//   * a seat's own per-suit HCP range is dropped when no later def exists (the `,@(if suit-hcp ..)`
//     splice of bridge.cl:1227-1229);
//   * lower bounds are inferred only when the four seats are all fixed or ranged
//     (hands-in-supply = others + 1, bridge.cl:1128,1194).
// One thread produces one deal: per ranged seat a rejection loop over uniform 13-subsets of the
// remaining cards (binary deal on compact positions + software deposit), then `deal-all` for the
// unranged seats.  This path serves the REST-sized jobs of the reference (hundreds of deals);
// joint rejection (the other kernels) is the throughput path.
#pragma once
#include <cstdint>
#include "deal.cuh"
#include "seatfirst.cuh"
#include "distgen.cuh"

namespace bmc {

constexpr uint32_t STREAM_REFORDER = 16u, STREAM_REFORDER_REST = 31u;
constexpr uint32_t SEQ_MAX_ATTEMPTS = 1u << 14;     // candidates per seat before the deal is given up (counted in `voided`)
constexpr uint32_t SEQ_DRAWS = 4u;                  // candidates a lane draws per pass of the pipeline

struct SeqSuit {
    int8_t has_spec, lo, hi, has_third, third_lo, third_hi;      // own spec; -1 = nil
    int8_t others_missing;                                        // some later def lacks this suit
    int8_t pad;
    int16_t sum_upper, sum_lower, sum_lower_third;                // over the later defs
    int16_t pad2;
};
struct SeqStage {
    int8_t has_hcp, hcp_lo, hcp_hi, others_missing_hcp, has_others, seat, pad[2];
    int16_t sum_upper_hcp, sum_lower_hcp;
    uint32_t m;                   // cards remaining when this seat is sampled
    SeqSuit suit[4];
    uint32_t thresh[13];          // Lemire thresholds of the 4-digit words of the 13-of-m binary deal
};
struct SeqDev {
    uint32_t n_stages, exhaust;   // exhaust: fixed + ranged seats = 4
    uint32_t q_rest;              // initial Q of the final deal-all (capacity 13 for the unranged seats)
    uint32_t n_rest;              // cards dealt by deal-all
    uint32_t thresh_rest[10];
    SeqStage st[4];
};

struct SeqRanges { int hcp_lo, hcp_hi; int len_lo[4], len_hi[4]; int sh_lo[4], sh_hi[4]; bool bad; };

__host__ __device__ __forceinline__ int seq_popc(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
__host__ __device__ __forceinline__ int seq_min(int a, int b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int seq_max(int a, int b) { return a > b ? a : b; }

// infer-distgen-params for stage `s` on the remaining cards (f0, f1: lane-layout masks)
__host__ __device__ __forceinline__ SeqRanges seq_infer(const SeqStage &s, uint32_t exhaust, uint32_t f0, uint32_t f1)
{
    SeqRanges r;
    r.bad = false;
    int slen[4], nlow[4], shcp[4], total = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t lane = ((k < 2 ? f0 : f1) >> (16 * (k & 1))) & 0x1FFFu;
        slen[k] = seq_popc(lane);
        nlow[k] = seq_popc(lane & 0x1FFu);
        shcp[k] = (int)((lane >> 9) & 1u) + 2 * (int)((lane >> 10) & 1u) + 3 * (int)((lane >> 11) & 1u) + 4 * (int)((lane >> 12) & 1u);
        total += shcp[k];
    }
    // ---- infer-hcp-range (bridge.cl:1126-1153)
    const int min_hcp = (exhaust && !s.others_missing_hcp) ? seq_max(total - (int)s.sum_upper_hcp, 0) : 0;
    const int up_limit = seq_min(total - (int)s.sum_lower_hcp, total);
    if (!s.has_hcp) {
        if (up_limit < total || min_hcp > 0) { r.hcp_lo = min_hcp; r.hcp_hi = up_limit; }
        else { r.hcp_lo = 0; r.hcp_hi = 64; }                      // no :hcp key: no filter
    } else {
        r.hcp_lo = seq_max(s.hcp_lo >= 0 ? (int)s.hcp_lo : 0, min_hcp);
        r.hcp_hi = s.hcp_hi >= 0 ? seq_min((int)s.hcp_hi, up_limit) : up_limit;
        if (r.hcp_lo > r.hcp_hi) r.bad = true;
    }
    // ---- infer-suit-ranges (bridge.cl:1192-1234)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const SeqSuit &q = s.suit[k];
        const int min_len = (exhaust && !q.others_missing) ? seq_max(slen[k] - (int)q.sum_upper, 0) : 0;
        const int ul = seq_min(slen[k] - (int)q.sum_lower, slen[k]);
        r.sh_lo[k] = 0; r.sh_hi[k] = 64;
        if (!q.has_spec) {
            if (ul < slen[k]) { r.len_lo[k] = min_len; r.len_hi[k] = ul; }
            else { r.len_lo[k] = 0; r.len_hi[k] = nlow[k]; }       // no key: the low-card cap
        } else {
            const int down = q.lo >= 0 ? seq_max((int)q.lo, min_len) : min_len;
            const int up = q.hi >= 0 ? seq_min((int)q.hi, ul) : ul;
            if (down > ul) r.bad = true;
            r.len_lo[k] = down; r.len_hi[k] = up;
            if (s.has_others) {                                      // suit-HCP narrowing only when later defs exist
                const int ulh = seq_min(shcp[k] - (int)q.sum_lower_third, shcp[k]);
                if (!q.has_third) {
                    if (ulh < shcp[k]) { r.sh_lo[k] = 0; r.sh_hi[k] = ulh; }
                } else {
                    r.sh_lo[k] = seq_max(q.third_lo >= 0 ? (int)q.third_lo : 0, 0);
                    r.sh_hi[k] = q.third_hi >= 0 ? seq_min((int)q.third_hi, ulh) : ulh;
                    if (r.sh_lo[k] > r.sh_hi[k]) r.bad = true;
                }
            }
        }
    }
    // Necessary conditions for a non-empty admissible set; where they fail the reference's generator has
    // total weight 0 and `(random 0)` signals an error (e.g. a 13-card single-suit remainder under the cap).
    int sum_lo = 0, sum_hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int hi = seq_min(r.len_hi[k], slen[k]), lo = seq_max(r.len_lo[k], 0);
        if (lo > hi) r.bad = true;
        sum_lo += lo;
        sum_hi += hi;
    }
    if (sum_lo > 13 || sum_hi < 13 || r.hcp_lo > total || r.hcp_hi < 0) r.bad = true;
    return r;
}

__device__ __forceinline__ bool seq_hand_ok(const SeqRanges &r, uint32_t m0, uint32_t m1)
{
    const Feat f = hand_features(m0, m1);
    const int hcp = (int)(f.f2 & 0xFFu);
    bool ok = hcp >= r.hcp_lo && hcp <= r.hcp_hi;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int len = (int)((f.f0 >> (8 * k)) & 0xFFu), sh = (int)((f.f1 >> (8 * k)) & 0xFFu);
        ok = ok && len >= r.len_lo[k] && len <= r.len_hi[k] && sh >= r.sh_lo[k] && sh <= r.sh_hi[k];
    }
    return ok;
}

// The narrowed ranges as the packed byte test of features.cuh (A = 0x80 - lo, B = 0x7F - hi per byte).
struct SeqPacked { uint32_t A0, B0, A1, B1, A2, B2; };
__host__ __device__ __forceinline__ SeqPacked seq_pack(const SeqRanges &r)
{
    SeqPacked p;
    p.A0 = p.B0 = p.A1 = p.B1 = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int llo = seq_max(r.len_lo[k], 0), lhi = seq_min(r.len_hi[k], 13);
        const int slo = seq_max(r.sh_lo[k], 0), shi = seq_min(r.sh_hi[k], 10);
        p.A0 |= (uint32_t)(0x80 - llo) << (8 * k); p.B0 |= (uint32_t)(0x7F - seq_max(lhi, 0)) << (8 * k);
        p.A1 |= (uint32_t)(0x80 - slo) << (8 * k); p.B1 |= (uint32_t)(0x7F - seq_max(shi, 0)) << (8 * k);
    }
    const int hlo = seq_max(r.hcp_lo, 0), hhi = seq_max(seq_min(r.hcp_hi, 37), 0);
    p.A2 = (uint32_t)(0x80 - hlo) | 0x80808000u;               // losers, balanced?, spare byte: unconstrained
    p.B2 = (uint32_t)(0x7F - hhi) | 0x7F7E0000u;
    return p;
}
__device__ __forceinline__ bool seq_packed_ok(const SeqPacked &p, uint32_t m0, uint32_t m1)
{
    const Feat f = hand_features(m0, m1);
    return in_ranges(f.f0, p.A0, p.B0) && in_ranges(f.f1, p.A1, p.B1) && in_ranges(f.f2, p.A2, p.B2);
}

// Stage 1 BY RANK: the first ranged seat goes through the root generator, whose narrowed ranges depend on the
// fixed hands only (bridge.cl:1297-1301 caches it for that reason).  Its admissible hands are enumerated once on
// the device (distgen.cuh: classes over the remaining deck, weights prod C(n_low_s, l_s)) and a hand is drawn by
// unranking a uniform R in [0, W) -- what `rand-hand` does with its weighted walk, without the walk.
struct SeqRoot {
    ExactTab x;
    unsigned long long W, S, S_recip, limit;     // hands; floor(2^64 / W); floor(2^64 / S); W S (u >= limit is redrawn)
    uint32_t low0, low1;                          // available low cards (lane layout, bits 0..8 of each lane)
    uint32_t full;                                // 1: every low card is available (no deposit needed)
};

// t-th l-subset of the first n of nine positions = the t-th entry of subset9's size-l list (numeric order puts the
// subsets that avoid the high positions first).
__device__ __forceinline__ bool seq_stage1(const PhiloxKey &key, const SeqRoot &root, const LowTabs &T, uint32_t idx_lo, uint32_t idx_hi,
                                           uint32_t &m0, uint32_t &m1)
{
    m0 = m1 = 0u;
    if (root.W == 0ull) return false;
    unsigned long long u = 0ull;
    bool got = false;
    for (uint32_t att = 0; att < 8u && !got; ++att) {           // u >= W S has probability < W / 2^64
        uint32_t w[4];
        philox4x32_10(key, idx_lo, idx_hi, att, STREAM_REFORDER, w);
        u = ((unsigned long long)w[0] << 32) | w[1];
        got = u < root.limit;
    }
    if (!got) return false;
    unsigned long long R = sf_mulhi64(u, root.S_recip);
    unsigned long long rem = u - R * root.S;
    while (rem >= root.S) { rem -= root.S; ++R; }
    const ExactTab &x = root.x;
    const uint32_t b = min(sf_mulhi32((uint32_t)(u >> 32), x.bmul), x.n_bkt - 1u);
    uint32_t lo = __ldg(x.bkt + b), hi = __ldg(x.bkt + b + 1u);
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1u) >> 1;
        if (__ldg(x.cum + mid) <= R) lo = mid; else hi = mid - 1u;
    }
    const uint32_t k = __ldg(x.keys + lo);
    uint32_t r = (uint32_t)(R - __ldg(x.cum + lo));
    uint32_t sub[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t l = min((k >> (16 + 4 * s)) & 15u, 9u);
        const uint32_t n = (uint32_t)__popc((((s < 2) ? root.low0 : root.low1) >> (16 * (s & 1))) & 0x1FFu);
        const uint32_t c = c_binom.c[n][l];                    // l-subsets of the n available low cards
        const uint32_t q = c ? r / c : 0u;
        sub[s] = T.subset9[(T.c9[l][2] + (r - q * c)) & 511u];
        r = q;
    }
    uint32_t l0 = sub[0] | (sub[1] << 16), l1 = sub[2] | (sub[3] << 16);
    if (!root.full) {
        // compact subsets -> the available positions: one deposit per word (the diamonds' / spades' bits follow the
        // clubs' / hearts' in the compact word)
        const uint32_t nc = (uint32_t)__popc(root.low0 & 0x1FFu), nh = (uint32_t)__popc(root.low1 & 0x1FFu);
        l0 = deposit_apply(deposit_prepare(root.low0), sub[0] | (sub[1] << nc));
        l1 = deposit_apply(deposit_prepare(root.low1), sub[2] | (sub[3] << nh));
    }
    m0 = l0 | ((k & 0xFu) << 9) | ((k & 0xF0u) << 21);
    m1 = l1 | ((k & 0xF00u) << 1) | ((k & 0xF000u) << 13);
    return true;
}

// Uniform 13-subset of the m <= 39 remaining cards as a mask over COMPACT positions (the free cards in card order),
// attempt `att` of stage stream `stream`: Floyd's algorithm -- for j = m-13 .. m-1 draw t uniform on [0, j] and take
// t, or j if t is taken already -- thirteen bounded draws, the mixed-radix digits (4 + 4 + 4 + 1) of the four words of
// ONE Philox call, each word with Lemire's remainder test (thresh[0..3] = 2^32 mod the product of its ranges).
// Returns false if a test failed (the caller simply moves on to the next attempt).
__device__ __forceinline__ bool seq_draw_floyd(const PhiloxKey &key, uint32_t idx_lo, uint32_t idx_hi, uint32_t stream, uint32_t att,
                                               uint32_t m, const uint32_t *__restrict__ thresh, unsigned long long &c)
{
    uint32_t w[4];
    philox4x32_10(key, idx_lo, idx_hi, att, stream, w);
    bool ok = true;
    c = 0ull;
    uint32_t j = m - 13u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t x = w[g];
#pragma unroll
        for (int d = 0; d < (g < 3 ? 4 : 1); ++d, ++j) {
            uint32_t t;
            mul_wide(x, j + 1u, x, t);
            const unsigned long long taken = (c >> t) & 1ull;
            c |= 1ull << (taken ? j : t);
        }
        ok = ok && !(x < thresh[g]);
    }
    return ok;
}

// Honour cards of the remaining deck in compact coordinates, as three weight planes: HCP of a compact subset c =
// popc(c & M1) + 2 popc(c & M2) + 4 popc(c & M4) (J = 1: M1; Q = 2: M2; K = 3: M1 and M2; A = 4: M4).
__device__ __forceinline__ void seq_hcp_planes(uint32_t f0, uint32_t f1, unsigned long long &M1, unsigned long long &M2, unsigned long long &M4)
{
    M1 = M2 = M4 = 0ull;
    const uint32_t base1 = (uint32_t)__popc(f0);
#pragma unroll
    for (int wd = 0; wd < 2; ++wd) {
        const uint32_t f = wd ? f1 : f0, base = wd ? base1 : 0u;
#pragma unroll
        for (int ln = 0; ln < 2; ++ln)
#pragma unroll
            for (int lev = 0; lev < 4; ++lev) {
                const uint32_t b = 16u * ln + 9u + lev;
                if ((f >> b) & 1u) {
                    const unsigned long long bit = 1ull << (base + (uint32_t)__popc(f & ((1u << b) - 1u)));
                    if (lev == 0 || lev == 2) M1 |= bit;
                    if (lev == 1 || lev == 2) M2 |= bit;
                    if (lev == 3) M4 |= bit;
                }
            }
    }
}
__device__ __forceinline__ int seq_compact_hcp(unsigned long long c, unsigned long long M1, unsigned long long M2, unsigned long long M4)
{
    return __popcll(c & M1) + 2 * __popcll(c & M2) + 4 * __popcll(c & M4);
}

// compact subset -> the hand in lane layout (software deposit into the free positions)
__device__ __forceinline__ void seq_expand(unsigned long long c, uint32_t f0, uint32_t f1, uint32_t &h0, uint32_t &h1)
{
    const uint32_t n0 = __popc(f0);
    const uint32_t a = n0 >= 32u ? (uint32_t)c : ((uint32_t)c & ((1u << n0) - 1u));
    const uint32_t b = (uint32_t)(c >> n0);
    h0 = deposit_apply(deposit_prepare(f0), a);
    h1 = deposit_apply(deposit_prepare(f1), b);
}

// One rejection attempt in full: draw, HCP in compact space (most candidates die here, cheaply), deposit, packed test.
__device__ __forceinline__ bool seq_attempt(const PhiloxKey &key, uint32_t idx_lo, uint32_t idx_hi, uint32_t stream, uint32_t att, uint32_t m,
                                            const uint32_t *__restrict__ thresh, uint32_t f0, uint32_t f1, const SeqPacked &pk, int hcp_lo,
                                            int hcp_hi, unsigned long long M1, unsigned long long M2, unsigned long long M4, uint32_t &h0, uint32_t &h1)
{
    unsigned long long c;
    if (!seq_draw_floyd(key, idx_lo, idx_hi, stream, att, m, thresh, c)) return false;
    const int hcp = seq_compact_hcp(c, M1, M2, M4);
    if (hcp < hcp_lo || hcp > hcp_hi) return false;
    seq_expand(c, f0, f1, h0, h1);
    return seq_packed_ok(pk, h0, h1);
}

}  // namespace bmc
