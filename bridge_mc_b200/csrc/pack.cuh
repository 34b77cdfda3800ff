// pack.cuh -- packed index records: about ONE byte per accepted deal on the way to the host.
//
// A deal is a pure function of (seed, index, mode, primary window), so an accepted deal can leave the device as its
// Philox index alone (bmc_sample_indices: a 32-bit offset, 4 bytes).  On a box where eight GPUs share the host's
// ingest bandwidth even that is the bottleneck (profiles/r02_bench_n8_v1.json: 93 GB/s for all ranks).  Accepted indices
// of a wave are a sparse subset of a contiguous range -- a bitmap -- and the gaps between neighbours are short
// (1 / acceptance on average), so each wave's offsets are scattered into a bitmap and the bitmap is written out as gap
// bytes, tile by tile:
//
//   chunk   = header (8 bytes, little endian: u32 offset of the tile's first index from the call's first index,
//             u32 payload length) + payload
//   payload = with cur = tile offset - 1: a byte b < 255 accepts index cur += b + 1; b = 255 skips: cur += 255
//
// Chunks are self-describing, so the blocks append them with one atomicAdd each, in any order; bmc_unpack_indices (host)
// turns the stream back into 32-bit offsets.  1.03 bytes per deal at an acceptance of 3.8 %.
#pragma once
#include <cstdint>

namespace bmc {

constexpr uint32_t PACK_TILE_BITS = 65536u, PACK_THREADS = 256u, PACK_WORDS = 8u;     // 256 threads x 8 words x 32 bits

struct PackState {
    unsigned long long prev_acc;      // accepted counter up to which offsets were packed
    unsigned long long bytes;         // byte cursor (keeps counting past the capacity)
    unsigned long long overflow;      // chunks that did not fit
    unsigned long long pad;
};

// offsets[prev_acc .. accepted) of the wave that just ran -> bits of the wave's bitmap
__global__ void pack_scatter_kernel(const uint32_t *__restrict__ offsets, const unsigned long long *__restrict__ tallies,
                                    const PackState *__restrict__ st, unsigned long long cap_offsets, uint32_t wave_base,
                                    uint32_t wave_n, uint32_t *__restrict__ bitmap)
{
    const unsigned long long from = st->prev_acc;
    unsigned long long to = tallies[2];
    if (to > cap_offsets) to = cap_offsets;
    for (unsigned long long j = from + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < to;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t off = offsets[j] - wave_base;
        if (off < wave_n) atomicOr(&bitmap[off >> 5], 1u << (off & 31u));
    }
}

__global__ void pack_advance_kernel(const unsigned long long *__restrict__ tallies, PackState *st, unsigned long long cap_offsets)
{
    const unsigned long long to = tallies[2];
    st->prev_acc = to > cap_offsets ? cap_offsets : to;
}

// one block per tile of 65 536 indices
__global__ void __launch_bounds__(PACK_THREADS) pack_tiles_kernel(const uint32_t *__restrict__ bitmap, uint32_t wave_base, uint32_t wave_n,
                                                                 uint8_t *__restrict__ out, unsigned long long cap_bytes, PackState *st)
{
    __shared__ int s_last[PACK_THREADS / 32];
    __shared__ uint32_t s_sum[PACK_THREADS / 32];
    __shared__ unsigned long long s_pos;
    __shared__ int s_ok;
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t tile_bit = blockIdx.x * PACK_TILE_BITS;
    const uint32_t n_words = (wave_n + 31u) >> 5;
    uint32_t w[PACK_WORDS];
    int last = -1;
#pragma unroll
    for (uint32_t k = 0; k < PACK_WORDS; ++k) {
        const uint32_t wi = (tile_bit >> 5) + t * PACK_WORDS + k;
        w[k] = wi < n_words ? bitmap[wi] : 0u;
        if (w[k]) last = (int)(t * 256u + k * 32u + (31u - (uint32_t)__clz((int)w[k])));
    }
    // exclusive max-scan of `last` over the threads: the accepted index before this thread's first one
    int incl = last;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl = max(incl, v); }
    if (lane == 31u) s_last[warp] = incl;
    __syncthreads();
    int prev = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0u) prev = -1;
    for (uint32_t k = 0; k < warp; ++k) prev = max(prev, s_last[k]);
    // bytes of this thread's gaps
    uint32_t nb = 0;
    {
        int p0 = prev;
#pragma unroll
        for (uint32_t k = 0; k < PACK_WORDS; ++k) {
            uint32_t m = w[k];
            while (m) {
                const int p = (int)(t * 256u + k * 32u + (uint32_t)__ffs((int)m) - 1u);
                m &= m - 1u;
                nb += 1u + (uint32_t)(p - p0 - 1) / 255u;
                p0 = p;
            }
        }
    }
    uint32_t scan = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, scan, o); if (lane >= (uint32_t)o) scan += v; }
    if (lane == 31u) s_sum[warp] = scan;
    __syncthreads();
    uint32_t off = scan - nb, total = 0;
    for (uint32_t k = 0; k < PACK_THREADS / 32; ++k) { if (k < warp) off += s_sum[k]; total += s_sum[k]; }
    if (total == 0u) return;
    if (t == 0u) {
        const unsigned long long pos = atomicAdd(&st->bytes, 8ull + total);
        const int ok = pos + 8ull + total <= cap_bytes;
        if (!ok) atomicAdd(&st->overflow, 1ull);
        else {
            const uint32_t base = wave_base + tile_bit;
            for (int i = 0; i < 4; ++i) { out[pos + i] = (uint8_t)(base >> (8 * i)); out[pos + 4 + i] = (uint8_t)(total >> (8 * i)); }
        }
        s_pos = pos; s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return;
    uint8_t *dst = out + s_pos + 8ull + off;
    int p0 = prev;
#pragma unroll
    for (uint32_t k = 0; k < PACK_WORDS; ++k) {
        uint32_t m = w[k];
        while (m) {
            const int p = (int)(t * 256u + k * 32u + (uint32_t)__ffs((int)m) - 1u);
            m &= m - 1u;
            uint32_t g = (uint32_t)(p - p0 - 1);               // gap - 1
            while (g >= 255u) { *dst++ = 255u; g -= 255u; }
            *dst++ = (uint8_t)g;
            p0 = p;
        }
    }
}

}  // namespace bmc
