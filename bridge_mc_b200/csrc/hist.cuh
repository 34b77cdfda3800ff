// hist.cuh -- device-side `hist-base` over measure tuples.
//
// The reference tallies the rows `(m1 m2 ...)` of an mc-case with hist-base (bridge.cl:1325-1330): equal
// tuples are counted and listed in first-seen order; hist-breakdown / hist-percent then build the tree the
// REST API returns (bridge.cl:1344-1412, rest.cl:276-277).  Here the measures that do not need the trick
// estimator are evaluated per deal on the device (one byte each, up to eight per tuple), the tuples are
// radix-sorted with their deal index (CUB, stable), run-length encoded, and the groups are returned in
// first-seen order -- exactly hist-base's output for the same rows.
#pragma once
#include <cub/cub.cuh>

#include "features.cuh"

namespace bmc {

enum MeasureKind : uint8_t {
    M_HCP = 0,          // HCP of `seat`
    M_LEN = 1,          // length of `suit` in `seat`
    M_SUIT_HCP = 2,     // HCP of `suit` in `seat`
    M_LOSERS = 3,       // losing-trick count of `seat` (bidding.cl:20-23)
    M_BALANCED = 4,     // balanced? of `seat`
    M_LINE_HCP = 5,     // HCP of `seat` and its partner (line-hcp, bridge.cl:3231-3236)
    M_TRUMP_LEN = 6,    // cards of `suit` held by `seat` and its partner (trump-length, bridge.cl:3083-3086)
    M_SHAPE = 7,        // sorted-shape class of `seat` (suit-distribution, bridge.cl:1438-1439; see bmc_shape_table)
    M_LONGEST = 8,      // longest combined suit of `seat` and its partner, later suit wins ties (longest-suit, bridge.cl:3030-3036)
};

struct MeasureDev { uint8_t kind[8], seat[8], suit[8]; uint32_t k; };

__global__ void __launch_bounds__(256) measure_kernel(const __grid_constant__ MeasureDev m, const uint4 *__restrict__ rec,
                                                      unsigned long long n, const uint8_t *__restrict__ shape_lut,
                                                      unsigned long long *__restrict__ keys, uint32_t *__restrict__ idx)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = rec[i];
    Feat f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t xl = (k & 1) ? 0u : 0xFFFFFFFFu, xh = (k & 2) ? 0u : 0xFFFFFFFFu;
        f[k] = hand_features((r.x ^ xl) & (r.z ^ xh) & 0x1FFF1FFFu, (r.y ^ xl) & (r.w ^ xh) & 0x1FFF1FFFu);
    }
    unsigned long long key = 0;
    for (uint32_t j = 0; j < m.k; ++j) {
        const uint32_t s = m.seat[j] & 3u, p = (s + 2u) & 3u, su = m.suit[j] & 3u;
        // select the two hands without dynamic register indexing
        const Feat a = s == 0u ? f[0] : s == 1u ? f[1] : s == 2u ? f[2] : f[3];
        const Feat b = p == 0u ? f[0] : p == 1u ? f[1] : p == 2u ? f[2] : f[3];
        uint32_t v = 0;
        switch (m.kind[j]) {
        case M_HCP: v = a.f2 & 0xFFu; break;
        case M_LEN: v = (a.f0 >> (8u * su)) & 0xFFu; break;
        case M_SUIT_HCP: v = (a.f1 >> (8u * su)) & 0xFFu; break;
        case M_LOSERS: v = (a.f2 >> 8) & 0xFFu; break;
        case M_BALANCED: v = (a.f2 >> 16) & 0xFFu; break;
        case M_LINE_HCP: v = (a.f2 & 0xFFu) + (b.f2 & 0xFFu); break;
        case M_TRUMP_LEN: v = ((a.f0 >> (8u * su)) & 0xFFu) + ((b.f0 >> (8u * su)) & 0xFFu); break;
        case M_SHAPE: v = shape_lut[(a.f0 & 0xFFu) * 196u + ((a.f0 >> 8) & 0xFFu) * 14u + ((a.f0 >> 16) & 0xFFu)]; break;
        case M_LONGEST: {
            const uint32_t L = a.f0 + b.f0;                      // bytewise sums (<= 13 each)
            uint32_t best = 0, bl = L & 0xFFu;
#pragma unroll
            for (uint32_t q = 1; q < 4; ++q) {
                const uint32_t l = (L >> (8u * q)) & 0xFFu;
                if (l >= bl) { bl = l; best = q; }
            }
            v = best;
            break;
        }
        default: break;
        }
        key |= (unsigned long long)(v & 0xFFu) << (8u * (7u - j));      // first measure = most significant byte
    }
    keys[i] = key;
    idx[i] = (uint32_t)i;
}

}  // namespace bmc
