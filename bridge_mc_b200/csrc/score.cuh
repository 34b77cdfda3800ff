// score.cuh -- duplicate scoring, IMPs and the stand-in trick source (device + host).
//
// contract-score (bridge.cl:3045-3069), base-score (compscore.cl:94-100) and match-points
// (compscore.cl:123-131) as pure integer functions, shared by score_kernel / imps_kernel, the
// index-keyed drill tally (score_tally_kernel) and the tally that is FUSED into stage B of the
// sampling kernels (kernels.cuh, BMC_TALLY_SCORE): scoring the accepted deals of a constrained
// sample in the launch that draws them.
//
// The reference's trick source is `play-deal` (bridge.cl:1488-2984, SURVEY 8f N1 -- not in this
// library), so a stand-in with the same interface (deal -> tricks of every contract) supplies the
// column: tricks_of() maps a 128-bit Philox counter to eight EXACTLY uniform values 0..13 -- the
// mixed-radix digits of one 32-bit word (14^8 < 2^32), with Lemire's remainder test deciding whether
// the word is usable; rejected words (31 %) fall through to the next word / the next Philox block.
//   * fused form: counter = the deal's own record (lo0, lo1, hi0, hi1), key = seed ^ "tricks": the
//     tricks are a function of the deal alone, like a real estimator -- independent of the sampling
//     mode, of the index that produced the deal and of how indices are split over GPUs;
//   * index form (bmc_score_tally): counter = (index_lo, index_hi, block, 1), key = seed.
// Mirrored bit for bit by orc_tricks_of / orc_score_tally_* in oracle/bmc_oracle.c.
#pragma once
#include "deal.cuh"

namespace bmc {

constexpr int SCORE_MAX_CONTRACTS = 8, SCORE_MAX_PAIRS = 8;
constexpr int SCORE_TR_WORDS = SCORE_MAX_CONTRACTS * 14, SCORE_IMP_WORDS = SCORE_MAX_PAIRS * 49;
constexpr int SCORE_WORDS = SCORE_TR_WORDS + SCORE_IMP_WORDS + SCORE_MAX_PAIRS;     // tricks | imps | dropped
constexpr unsigned long long SCORE_KEY_XOR = 0x0000747269636B73ull;                  // "tricks"

struct ContractDev { int8_t level, suit, declarer, dbl; };      // same layout as bmc_contract
struct ContractTable {
    ContractDev c[SCORE_MAX_CONTRACTS];
    uint8_t vul[SCORE_MAX_CONTRACTS];
    uint8_t pairs[2 * SCORE_MAX_PAIRS];
    uint32_t n, n_pairs;
};

// contract-score, written out from the rules of the reference text (incl. its non-standard
// branches).  level 0 = not given: max(tricks - 6, 1).
__host__ __device__ inline int contract_score_dev(int suit, int tricks, int vul, int dbl, int level)
{
    if (level == 0) level = tricks - 6 > 1 ? tricks - 6 : 1;
    const int r = tricks - level - 6;
    const int mult = dbl == 2 ? 4 : (dbl ? 2 : 1);
    if (r < 0) {
        if (!dbl) return r * (vul ? 100 : 50);
        int pen;
        if (vul) pen = -200 + 300 * (r + 1);
        else pen = r >= -3 ? 100 + 200 * r : -800 + 300 * (r + 3);      // -100 -300 -500, then -1100 ...
        return pen * (dbl == 2 ? 2 : 1);
    }
    const int per = suit <= 1 ? 20 : 30;
    const int trick_pts = ((suit == 4 ? 10 : 0) + per * level) * mult;
    int bonus;
    if (level == 7) bonus = vul ? 2000 : 1300;
    else if (level == 6) bonus = vul ? 1250 : 800;
    else bonus = trick_pts >= 100 ? (vul ? 500 : 300) : 50;
    const int over = dbl ? (dbl == 2 ? 100 : 50) + (vul ? 200 : 100) * r : per * r;
    return trick_pts + bonus + over;
}
__host__ __device__ inline int base_score_dev(ContractDev c, int tricks, int vul)
{
    const int s = contract_score_dev(c.suit, tricks, vul, c.dbl, c.level);
    return (c.declarer == 1 || c.declarer == 3) ? -s : s;
}

// match-points: index of the first threshold > |diff|, signed; bad where the reference errors (|diff| >= 4000)
__host__ __device__ inline int imps_dev(int diff, bool &bad)
{
    constexpr int pts[24] = {20, 50, 90, 130, 170, 220, 270, 320, 370, 430, 500, 600, 750, 900,
                             1100, 1300, 1500, 1750, 2000, 2300, 2500, 3000, 3500, 4000};
    const int a = diff < 0 ? -diff : diff;
    int pos = 24;
#pragma unroll
    for (int i = 23; i >= 0; --i) if (a < pts[i]) pos = i;
    bad = pos == 24;
    return bad ? 0 : (diff < 0 ? -pos : pos);
}

// 14^8 = 1 475 789 056; 2^32 mod 14^8 = 1 343 389 184: a word whose final remainder is below it is rejected
constexpr uint32_t TRICK_LEMIRE = 1343389184u;
constexpr uint32_t TRICK_MAX_BLOCKS = 16u;          // 64 words: all rejected with probability 0.313^64 < 1e-32

// Eight exactly uniform trick counts from the Philox stream (c0, c1, block, c3) -- or, in the
// deal-keyed form, (c0, c1, c2 + block, c3).  t[j] = j-th mixed-radix digit (multiply-shift).
__device__ __forceinline__ void tricks_of(const PhiloxKey &key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t t[8])
{
    for (uint32_t b = 0;; ++b) {
        uint32_t w[4];
        philox4x32_10(key, c0, c1, c2 + b, c3, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t x = w[j], d[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) mul_wide(x, 14u, x, d[k]);
            if (x >= TRICK_LEMIRE || (b + 1u == TRICK_MAX_BLOCKS && j == 3)) {
#pragma unroll
                for (int k = 0; k < 8; ++k) t[k] = d[k];
                return;
            }
        }
    }
}

// One deal's contribution to the shared-memory score histogram `h` (SCORE_WORDS words):
// tricks_hist[c][t], imp_hist[p][24 + imps], dropped[p]; s_score = base-score table [8][14].
__device__ __forceinline__ void score_tally_update(const ContractTable &t, const int *s_score, const uint32_t tr[8], uint32_t *h)
{
#pragma unroll
    for (int c = 0; c < SCORE_MAX_CONTRACTS; ++c)
        if (c < (int)t.n) atomicAdd(&h[c * 14 + tr[c]], 1u);
#pragma unroll
    for (int p = 0; p < SCORE_MAX_PAIRS; ++p)
        if (p < (int)t.n_pairs) {
            const int ca = t.pairs[2 * p], cb = t.pairs[2 * p + 1];
            // select without dynamic register indexing
            uint32_t ta = 0, tb = 0;
#pragma unroll
            for (int c = 0; c < SCORE_MAX_CONTRACTS; ++c) { ta = ca == c ? tr[c] : ta; tb = cb == c ? tr[c] : tb; }
            bool bad;
            const int v = imps_dev(s_score[ca * 14 + ta] - s_score[cb * 14 + tb], bad);
            atomicAdd(bad ? &h[SCORE_TR_WORDS + SCORE_IMP_WORDS + p] : &h[SCORE_TR_WORDS + p * 49 + 24 + v], 1u);
        }
}

}  // namespace bmc
