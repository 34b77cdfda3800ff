// filter.cuh -- K2: features / accept flag / choose-bid on GIVEN deals (device code only).
//
// assess-hand, balanced? (bidding.cl:33-46), test-bid / choose-bid (bidding.cl:111-112, 146-150) over a deal
// set: one thread per record, four seats each.  HBM-bound when the rule table is cheap to evaluate:
// 16 B in + 49 B out per deal.
//
// Compiled twice, like kernels.cuh: ahead of time (the table's rows are interpreted from bytecode, one
// program after the other until one holds), and -- for large or repeated jobs -- at run time by NVRTC with
// BMC_JIT_FILTER defined, where `bmc_jit_choose_bid` is C++ generated from the table's rule forms (jit.hpp):
// the rows become one straight-line function whose shared sub-expressions the compiler folds, the way the
// reference `eval`s (compiles) a rule form each time it tests one (good-opening?, bidding.cl:60-74).
#pragma once
#include "kernels.cuh"

struct SchemeDev {
    const uint32_t *code;
    const uint32_t *row_off;     // offsets into code, n_rows entries
    uint32_t n_rows;
};

// same layout as bmc_features (include/bridgemc.h)
struct SeatOut { uint8_t len[4], suit_hcp[4], hcp, losers, balanced; int8_t bid; };
struct FeatOut { SeatOut seat[4]; };

__device__ __forceinline__ void filter_body(const ConsDev &cons, int has_cons, const SchemeDev &sch, const uint4 *__restrict__ rec,
                                            unsigned long long n, uint8_t *__restrict__ accept, FeatOut *__restrict__ feats,
                                            uint32_t *s_feat /* [8][FEAT_SCRATCH_WORDS] */)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (live) r = rec[i];
    const Planes pl = {r.x, r.y, r.z, r.w};
    bool ok = true;
    FeatOut out;
    uint32_t *scratch = s_feat + (threadIdx.x >> 5) * FEAT_SCRATCH_WORDS;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t m0, m1;
        seat_words(pl, k, m0, m1);
        const Feat f = hand_features(m0, m1);
        const uint8_t *fb = feat_store(scratch, threadIdx.x & 31u, f);
        if (has_cons) {
            const SeatC &c = cons.seat[k];
            if (c.active) {
                ok = ok && seat_ranges_ok(c, f);
                if (c.prog != 0xFFFFFFFFu) {
                    bool err = false;
                    const bool v = run_program(cons.code + c.prog, fb, err);
                    ok = ok && v && !err;
                }
            }
        }
        int bid = -1;
#ifdef BMC_JIT_FILTER
        bid = bmc_jit_choose_bid(f);
#else
        if (sch.n_rows) {
            // choose-bid: first row whose form is true; -2 where the reference would
            // signal an error before finding one
            bool done = false;
            for (uint32_t row = 0; row < sch.n_rows; ++row) {
                bool err = false;
                const bool v = run_program(sch.code + __ldg(sch.row_off + row), fb, err);
                if (!done) {
                    if (err) { bid = -2; done = true; }
                    else if (v) { bid = (int)row; done = true; }
                }
                if (__all_sync(0xFFFFFFFFu, done)) break;
            }
        }
#endif
        SeatOut &o = out.seat[k];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            o.len[s] = (uint8_t)((f.f0 >> (8 * s)) & 0xFFu);
            o.suit_hcp[s] = (uint8_t)((f.f1 >> (8 * s)) & 0xFFu);
        }
        o.hcp = (uint8_t)(f.f2 & 0xFFu);
        o.losers = (uint8_t)((f.f2 >> 8) & 0xFFu);
        o.balanced = (uint8_t)((f.f2 >> 16) & 0xFFu);
        o.bid = (int8_t)bid;
    }
    if (!live) return;
    if (accept) accept[i] = ok ? 1 : 0;
    if (feats) {
        const uint4 *src = reinterpret_cast<const uint4 *>(&out);
        uint4 *dst = reinterpret_cast<uint4 *>(feats + i);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
    }
}

#ifndef BMC_JIT
__global__ void __launch_bounds__(256) filter_kernel(const __grid_constant__ ConsDev cons, int has_cons,
                                                     const __grid_constant__ SchemeDev sch, const uint4 *__restrict__ rec,
                                                     unsigned long long n, uint8_t *__restrict__ accept, FeatOut *__restrict__ feats)
{
    __shared__ uint32_t s_feat[8 * FEAT_SCRATCH_WORDS];
    filter_body(cons, has_cons, sch, rec, n, accept, feats, s_feat);
}
#elif defined(BMC_JIT_FILTER)
extern "C" __global__ void __launch_bounds__(256) bmc_jit_filter(const __grid_constant__ ConsDev cons, int has_cons,
                                                                 const __grid_constant__ SchemeDev sch, const uint4 *__restrict__ rec,
                                                                 unsigned long long n, uint8_t *__restrict__ accept, FeatOut *__restrict__ feats)
{
    __shared__ uint32_t s_feat[8 * FEAT_SCRATCH_WORDS];
    filter_body(cons, has_cons, sch, rec, n, accept, feats, s_feat);
}
#endif
