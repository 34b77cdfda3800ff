// distgen.cuh -- the reference's exact hand enumerator on the device (bridge.cl:866-1045).
//
// `distgen` walks every honour pattern -- the subsets of the available J Q K A x c d h s, binary counting
// order with the club jack most significant (hc-permuts, bridge.cl:921-963) -- and weights it by the
// number of low-card completions that meet the suit-length ranges: possible-lc-lens x permut-weight
// (bridge.cl:990-1004, 828-831), weight = sum over admissible (l_c l_d l_h l_s) of prod C(n_low_s, l_s).
// Here one thread block takes one pattern and one thread one low-card length vector: a CLASS of hands
// (honours exactly, low-card counts per suit), all of whose eleven features -- lengths, per-suit HCP,
// HCP, losers, balanced? -- are fixed.  Two consumers:
//   * bmc_distgen: the (weight pattern) rows, total weight and exact HCP / suit-length pmfs of a range
//     spec over any remaining deck, bit-exact against backend/test/distgen-test.dat;
//   * the exact class table of the seat-first sampler (kernels.cuh, mode 4): every class the primary
//     seat's constraint -- ranges AND rule program -- admits, with cumulative hand counts, so that stage
//     A1 accepts exactly the indices whose hand satisfies the seat and stage A2 unranks it directly.
#pragma once
#include "features.cuh"

namespace bmc {

constexpr int DG_THREADS = 1024;            // 10 x 10 x 10 low-card vectors (l_c l_d l_h), l_s implied
constexpr int DG_PATTERNS = 65536;

struct ClassSpec {
    SeatC seat;                   // byte-range test over the eleven features
    const uint32_t *prog;         // rule program of the seat, or nullptr
    uint32_t n_low[4];            // low cards (2..10) available per suit
    uint32_t hon_avail;           // available honours, pattern order: bit 15 = club J, 14 = club Q ... bit 0 = spade A
    uint32_t hand_size;           // 13
};

// C(n, k) for n, k <= 9
__host__ __device__ constexpr uint32_t dg_binom(uint32_t n, uint32_t k)
{
    if (k > n) return 0u;
    uint32_t r = 1u;
    for (uint32_t i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}
struct BinomTab { uint32_t c[10][10]; };
__host__ __device__ constexpr BinomTab make_binom_tab()
{
    BinomTab t{};
    for (uint32_t n = 0; n < 10; ++n)
        for (uint32_t k = 0; k < 10; ++k) t.c[n][k] = dg_binom(n, k);
    return t;
}
__constant__ BinomTab c_binom = make_binom_tab();

// pattern (bit 15 = club J ... bit 0 = spade A) -> honour nibbles in lane order (bit 4 s + level, J = 0 .. A = 3)
__host__ __device__ __forceinline__ uint32_t dg_pattern_to_h(uint32_t pat)
{
    uint32_t h = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) h |= ((pat >> (15 - i)) & 1u) << i;
    return h;
}

struct ClassEval { bool pattern_ok, admitted; uint32_t key; uint32_t weight; Feat f; };

// The class (pattern, l0, l1, l2): a representative hand (the honours + the lowest l_s low cards of each
// suit) goes through the SAME hand_features / range test / rule interpreter as a sampled hand.
__device__ __forceinline__ ClassEval dg_eval(const ClassSpec &cs, uint32_t pat, uint32_t l0, uint32_t l1, uint32_t l2,
                                             uint32_t *scratch, unsigned lane)
{
    ClassEval e;
    const uint32_t H = dg_pattern_to_h(pat);
    const int l3 = (int)cs.hand_size - (int)__popc(H) - (int)(l0 + l1 + l2);
    const bool valid = l0 <= cs.n_low[0] && l1 <= cs.n_low[1] && l2 <= cs.n_low[2] && l3 >= 0 && l3 <= (int)cs.n_low[3];
    const uint32_t l3u = valid ? (uint32_t)l3 : 0u;
    const uint32_t m0 = ((1u << l0) - 1u) | ((H & 0xFu) << 9) | ((((1u << l1) - 1u) | (((H >> 4) & 0xFu) << 9)) << 16);
    const uint32_t m1 = ((1u << l2) - 1u) | (((H >> 8) & 0xFu) << 9) | ((((1u << l3u) - 1u) | (((H >> 12) & 0xFu) << 9)) << 16);
    e.f = hand_features(m0, m1);
    // honour-level tests (hc-permuts' filter: total HCP and per-suit HCP): byte 0 of word 2, all of word 1
    e.pattern_ok = in_ranges(e.f.f1, cs.seat.A[1], cs.seat.B[1]) &&
                   ((((e.f.f2 & 0xFFu) + (cs.seat.A[2] & 0xFFu)) & ~((e.f.f2 & 0xFFu) + (cs.seat.B[2] & 0xFFu))) & 0x80u) != 0u;
    bool ok = valid && seat_ranges_ok(cs.seat, e.f);
    if (cs.prog != nullptr) {                                  // warp-uniform: every lane interprets (inactive lanes are ignored)
        bool err = false;
        const bool v = run_program(cs.prog, feat_store(scratch, lane, e.f), err);
        ok = ok && v && !err;
    }
    e.admitted = ok;
    e.weight = valid ? c_binom.c[cs.n_low[0]][min(l0, 9u)] * c_binom.c[cs.n_low[1]][min(l1, 9u)] * c_binom.c[cs.n_low[2]][min(l2, 9u)] *
                           c_binom.c[cs.n_low[3]][min(l3u, 9u)] : 0u;
    e.key = H | (l0 << 16) | (l1 << 20) | (l2 << 24) | (l3u << 28);
    return e;
}

// Pass 1: per pattern, the number of admitted classes and their total weight; optionally the exact pmfs.
// pat_flag[p] = 1 where the pattern passes hc-permuts' filter (it is a row of the generator even when no
// low-card vector fits it: weight 0).
__global__ void __launch_bounds__(DG_THREADS) distgen_count_kernel(const __grid_constant__ ClassSpec cs, unsigned long long *__restrict__ pat_w,
                                                                   uint32_t *__restrict__ pat_cnt, uint8_t *__restrict__ pat_flag,
                                                                   unsigned long long *__restrict__ hcp_pmf /*[40] or null*/,
                                                                   unsigned long long *__restrict__ len_pmf /*[4][14] or null*/)
{
    __shared__ uint32_t s_feat[DG_THREADS / 32][FEAT_SCRATCH_WORDS];
    __shared__ unsigned long long s_w[DG_THREADS / 32];
    __shared__ uint32_t s_c[DG_THREADS / 32];
    __shared__ unsigned long long s_len[4 * 14];
    const uint32_t pat = blockIdx.x;
    const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    if (pat & ~cs.hon_avail) {                                  // holds an honour that is not in the deck
        if (t == 0) { pat_w[pat] = 0ull; pat_cnt[pat] = 0u; pat_flag[pat] = 0; }
        return;
    }
    if (t < 56) s_len[t] = 0ull;
    __syncthreads();
    const bool in = t < 1000u;
    const ClassEval e = dg_eval(cs, pat, in ? t / 100u : 15u, in ? (t / 10u) % 10u : 15u, in ? t % 10u : 15u, s_feat[warp], lane);
    const bool adm = in && e.admitted && e.pattern_ok;
    unsigned long long w = adm ? (unsigned long long)e.weight : 0ull;
    uint32_t c = adm ? 1u : 0u;
    if (adm && len_pmf) {
#pragma unroll
        for (int s = 0; s < 4; ++s) atomicAdd(&s_len[s * 14 + ((e.f.f0 >> (8 * s)) & 0xFFu)], w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { w += __shfl_xor_sync(0xFFFFFFFFu, w, o); c += __shfl_xor_sync(0xFFFFFFFFu, c, o); }
    if (lane == 0) { s_w[warp] = w; s_c[warp] = c; }
    __syncthreads();
    if (t == 0) {
        unsigned long long tw = 0; uint32_t tc = 0;
        for (int i = 0; i < DG_THREADS / 32; ++i) { tw += s_w[i]; tc += s_c[i]; }
        pat_w[pat] = tw; pat_cnt[pat] = tc;
        // thread 0's class has l = (0 0 0): its pattern-level verdict is the pattern's
        pat_flag[pat] = e.pattern_ok ? 1 : 0;
        if (hcp_pmf && tw) atomicAdd(&hcp_pmf[e.f.f2 & 0xFFu], tw);
    }
    if (len_pmf && t < 56 && s_len[t]) atomicAdd(&len_pmf[t], s_len[t]);
}

// Pass 2: the admitted classes in (pattern, l_c, l_d, l_h) order: keys[i] and cum[i] = hands before class i.
__global__ void __launch_bounds__(DG_THREADS) distgen_emit_kernel(const __grid_constant__ ClassSpec cs, const unsigned long long *__restrict__ pat_cum,
                                                                  const uint32_t *__restrict__ pat_off, const uint32_t *__restrict__ pat_cnt,
                                                                  uint32_t *__restrict__ keys, unsigned long long *__restrict__ cum)
{
    __shared__ uint32_t s_feat[DG_THREADS / 32][FEAT_SCRATCH_WORDS];
    __shared__ unsigned long long s_w[DG_THREADS / 32];
    __shared__ uint32_t s_c[DG_THREADS / 32];
    const uint32_t pat = blockIdx.x;
    if (pat_cnt[pat] == 0u) return;                             // block-uniform
    const unsigned t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const bool in = t < 1000u;
    const ClassEval e = dg_eval(cs, pat, in ? t / 100u : 15u, in ? (t / 10u) % 10u : 15u, in ? t % 10u : 15u, s_feat[warp], lane);
    const bool adm = in && e.admitted && e.pattern_ok;
    // block-wide exclusive scan of (count, weight) in thread order
    unsigned long long w = adm ? (unsigned long long)e.weight : 0ull, wi = w;
    uint32_t c = adm ? 1u : 0u, ci = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long w2 = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        const uint32_t c2 = __shfl_up_sync(0xFFFFFFFFu, ci, o);
        if (lane >= (unsigned)o) { wi += w2; ci += c2; }
    }
    if (lane == 31) { s_w[warp] = wi; s_c[warp] = ci; }
    __syncthreads();
    unsigned long long wb = 0; uint32_t cb = 0;
    for (unsigned i = 0; i < warp; ++i) { wb += s_w[i]; cb += s_c[i]; }
    if (adm) {
        const uint32_t at = pat_off[pat] + cb + (ci - c);
        keys[at] = e.key;
        cum[at] = pat_cum[pat] + wb + (wi - w);
    }
}

// Bucket index of the class search: bkt[j] = the class holding the smallest u whose bucket umulhi(u_hi, bmul) is j
// (clamped to the last class); bkt[n_bkt] = n_cls - 1.  S = floor(2^64 / C(52,13)): rank of u = u / S.
__global__ void distgen_bucket_kernel(const unsigned long long *__restrict__ cum, uint32_t n_cls, uint32_t bmul, uint32_t n_bkt,
                                      unsigned long long S, uint32_t *__restrict__ bkt)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_bkt) return;
    if (j == n_bkt) { bkt[j] = n_cls - 1u; return; }
    // smallest u_hi with umulhi(u_hi, bmul) >= j
    unsigned long long lo = (((unsigned long long)j << 32) + bmul - 1ull) / bmul;
    while (lo > 0ull && (uint32_t)((lo - 1ull) * bmul >> 32) >= j) --lo;
    while (lo <= 0xFFFFFFFFull && (uint32_t)(lo * bmul >> 32) < j) ++lo;
    if (lo > 0xFFFFFFFFull) { bkt[j] = n_cls - 1u; return; }
    const unsigned long long R = (lo << 32) / S;                // rank of the smallest u of the bucket
    uint32_t a = 0, b = n_cls - 1u;                             // largest class with cum <= R
    while (a < b) {
        const uint32_t mid = (a + b + 1u) >> 1;
        if (cum[mid] <= R) a = mid; else b = mid - 1u;
    }
    bkt[j] = a;
}

}  // namespace bmc
