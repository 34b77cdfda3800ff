// playdeal_kernels.cu -- the kernels of the trick estimator (playdeal.cuh), in their own translation unit.
//
// The estimator is a long scalar search: some 10 000 SASS instructions with no hot loop.  ncu shows that with several deals
// resident per SM instruction fetch dominates (19.7 of 27.2 cycles per issued instruction stalled on `no_instruction`,
// profiles/r02_prof_playdeal_full.txt): the L1.5 instruction cache holds 32 KB.  This unit is therefore built for SIZE.
#include "playdeal.cuh"
#include "playdeal_host.h"
#include "../../include/bridgemc.h"

namespace {

__device__ __noinline__ void play_one_deal(const uint4 *records, const unsigned long long *hands4, unsigned long long i, const int8_t *trump,
                                           const uint8_t *declarer, int per_deal, uint8_t *arena, uint32_t arena_bytes, pd::Scratch &S,
                                           bmc_play_result *out, int timing, unsigned int *probe)
{
    const long long t0 = timing ? clock64() : 0ll;
    const int tr = trump[per_deal ? i : 0];
    pd::Hand h[4];
    if (records) {
        const uint4 r = records[i];                              // lo plane (x, y), hi plane (z, w)
        const int dec = declarer[per_deal ? i : 0] & 3;
        for (int k = 0; k < 4; ++k) {
            const int seat = (dec + k) & 3;
            const uint32_t xl = (seat & 1) ? 0u : 0xFFFFFFFFu, xh = (seat & 2) ? 0u : 0xFFFFFFFFu;
            const uint32_t m0 = (r.x ^ xl) & (r.z ^ xh) & 0x1FFF1FFFu, m1 = (r.y ^ xl) & (r.w ^ xh) & 0x1FFF1FFFu;
            h[k].s.m = (unsigned long long)m0 | ((unsigned long long)m1 << 32);      // the record's lane layout is the hand's
            h[k].set_id(k);
        }
    } else {
        for (int k = 0; k < 4; ++k) {                            // four lane masks: declarer, left, dummy, right
            h[k].s.m = hands4[4 * i + k] & 0x1FFF1FFF1FFF1FFFull;
            h[k].set_id(k);
        }
    }
    pd::Result &res = S.res;
    pd::play_deal(arena, arena_bytes, tr < 0 || tr > 3 ? -1 : tr, h, S, res, probe != nullptr);
    if ((threadIdx.x & 31u) != 0u) return;                   // the lanes hold identical results
    if (probe) { probe[i] = res.nodes; return; }
    bmc_play_result o;
    o.status = (uint8_t)res.status; o.n_tricks = res.n; o.score[0] = res.score[0]; o.score[1] = res.score[1];
    for (int t = 0; t < 13; ++t) {
        for (int c = 0; c < 4; ++c) o.cards[t][c] = t < res.n ? res.tricks[t][c] : 0xFF;
        o.winner[t] = t < res.n ? res.wno[t] : 0xFF;
    }
    // diagnostic (BMC_PD_TIMING=1): SM cycles spent on this deal, in units of 4096, 24 bits little endian; else zero
    const unsigned long long dt = timing ? (unsigned long long)(clock64() - t0) >> 12 : 0ull;
    const unsigned int d24 = dt > 0xFFFFFFull ? 0xFFFFFFu : (unsigned int)dt;
    o.reserved[0] = (uint8_t)d24; o.reserved[1] = (uint8_t)(d24 >> 8); o.reserved[2] = (uint8_t)(d24 >> 16);
    out[i] = o;
}

// One WARP per deal.  The estimator is a long scalar computation whose working set (search stacks, route lists) is a few
// tens of KB: with one deal per thread a warp drags 32 such working sets through the caches and serialises 32 divergent
// control flows.  Here the 32 lanes of a warp run the SAME deal in lock step on the same data -- the search stacks sit in
// shared memory, one copy per warp, the arena in global memory, one per warp -- so there is no divergence, every load is a
// broadcast and every store a single transaction.  Deals are handed out by an atomic counter (their cost varies 100-fold).
__global__ void __launch_bounds__(bmc_pd::THREADS) play_deal_kernel(const uint4 *records, const unsigned long long *hands4, const unsigned int *idx,
                                                                      unsigned long long n, const int8_t *trump, const uint8_t *declarer,
                                                                      int per_deal, uint8_t *arenas, uint32_t arena_bytes, bmc_play_result *out,
                                                                      unsigned long long *next, int timing, unsigned int *probe)
{
    extern __shared__ __align__(16) unsigned char pd_smem[];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    pd::Scratch &S = reinterpret_cast<pd::Scratch *>(pd_smem)[warp];
    uint8_t *arena = arenas + ((unsigned long long)blockIdx.x * bmc_pd::WARPS + warp) * (unsigned long long)arena_bytes;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(next, 1ull);
        k = __shfl_sync(0xFFFFFFFFu, k, 0);
        if (k >= n) break;
        play_one_deal(records, hands4, idx ? idx[k] : k, trump, declarer, per_deal, arena, arena_bytes, S, out, timing, probe);
        __syncwarp();
    }
}

// deals whose route lists outgrew their arena (status 1): their indices, compacted, for the retry pass
__global__ void play_deal_failed_kernel(const bmc_play_result *out, unsigned long long n, unsigned int *idx, unsigned int *count)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && out[i].status == 1) idx[atomicAdd(count, 1u)] = (unsigned int)i;
}

}  // namespace

namespace bmc_pd {

cudaError_t configure(int *blocks_per_sm)
{
    const size_t smem = WARPS * sizeof(pd::Scratch);
    cudaError_t e = cudaFuncSetAttribute(play_deal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, play_deal_kernel, THREADS, smem);
}

void launch(unsigned blocks, cudaStream_t stream, const void *d_records, const void *d_hands4, const unsigned int *d_idx, uint64_t n,
            const int8_t *d_trump, const uint8_t *d_declarer, int per_deal, uint8_t *d_arenas, uint32_t arena_bytes, void *d_out,
            unsigned long long *d_next, int timing, unsigned int *d_probe)
{
    play_deal_kernel<<<blocks, THREADS, WARPS * sizeof(pd::Scratch), stream>>>(
        reinterpret_cast<const uint4 *>(d_records), reinterpret_cast<const unsigned long long *>(d_hands4), d_idx, n, d_trump, d_declarer,
        per_deal, d_arenas, arena_bytes, reinterpret_cast<bmc_play_result *>(d_out), d_next, timing, d_probe);
}

void launch_failed(cudaStream_t stream, const void *d_out, uint64_t n, unsigned int *d_idx, unsigned int *d_count)
{
    play_deal_failed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const bmc_play_result *>(d_out), n, d_idx, d_count);
}

}  // namespace bmc_pd

#if defined(PD_PROFILE)
// diagnostic builds only: reads and clears the per-phase cycle sums
extern "C" __attribute__((visibility("default"))) int bmc_pd_debug_profile(unsigned long long *out8)
{
    unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out8, g_pd_prof, sizeof(zero)) != cudaSuccess) return 1;
    return cudaMemcpyToSymbol(g_pd_prof, zero, sizeof(zero)) != cudaSuccess;
}
#endif
