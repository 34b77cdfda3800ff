// bridgemc.cu -- kernels and C ABI of libbridgemc.so (sm_100a).
//
// Kernels
//   sample_kernel   K1: Philox deal + first-stage range test + warp queue +
//                   second-stage full test + tallies + compacted record stores
//   filter_kernel   K2: features / accept flag / choose-bid on given records
//   score_kernel    K4: base-score over (contract, tricks) tables
//   imps_kernel     K4: match-points
//   score_tally_kernel  K4 fused drill tally
//   int_peak_kernel INT32 issue-rate microbenchmark (roofline denominator)
// Nothing here is a contraction, so no tensor cores; the binding resource is
// integer issue.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/bridgemc.h"
#include "deal.cuh"
#include "features.cuh"
#include "seatfirst.cuh"
#include "reforder.cuh"
#include "hist.cuh"
#include "distgen.cuh"
#include "playdeal_host.h"
#include "pack.cuh"
#include "rules.hpp"
#include "jit.hpp"
#include "embedded_src.inc"

using namespace bmc;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(BMC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

#include "kernels.cuh"

// ---------------------------------------------------------------------------
// K1c: reference-order sampler (reforder.cuh): one thread = one deal, ranged seats sampled one
// after the other exactly as `build` does (bridge.cl:1292-1319).
// ---------------------------------------------------------------------------
// One thread, one index, start to end (bmc_expand; the sampling kernel below pipelines the same steps and draws the
// same numbers): stage 1 by rank, every later ranged seat by rejection over uniform 13-subsets of the remaining
// cards, deal-all for the rest.  Returns true where the reference signals an error for the index (`voided`).
__device__ __forceinline__ bool reforder_deal(const SampleArgs &a, const SeqDev &q, const SeqRoot &root, const LowTabs &T,
                                              unsigned long long idx, bool valid, Planes &pl)
{
    const uint32_t il = (uint32_t)idx, ih = (uint32_t)(idx >> 32);
    uint32_t lo0 = a.fix.lo0, lo1 = a.fix.lo1, hi0 = a.fix.hi0, hi1 = a.fix.hi1;
    uint32_t f0 = a.fix.free0, f1 = a.fix.free1;
    bool bad = false;
    for (uint32_t j = 0; j < q.n_stages; ++j) {
        const SeqStage &st = q.st[j];
        if (j > 0u && st.m <= 13u) break;                     // a LATER seat is implied once 13 cards remain (bridge.cl:1303);
                                                              // the first def always goes through the root generator
        uint32_t h0 = 0, h1 = 0;
        if (j == 0u) {
            if (valid && !seq_stage1(a.key, root, T, il, ih, h0, h1)) bad = true;
        } else {
            const SeqRanges r = seq_infer(st, q.exhaust, f0, f1);
            const SeqPacked pk = seq_pack(r);
            bad = bad || r.bad;
            unsigned long long M1, M2, M4;
            seq_hcp_planes(f0, f1, M1, M2, M4);
            uint32_t att = 0;
            bool need = valid && !bad;
            while (need) {
                if (seq_attempt(a.key, il, ih, STREAM_REFORDER + j, att, st.m, st.thresh, f0, f1, pk, r.hcp_lo, r.hcp_hi, M1, M2, M4, h0, h1)) need = false;
                else if (++att >= SEQ_MAX_ATTEMPTS) { bad = true; need = false; }
            }
        }
        if (!valid || bad) { h0 = 0; h1 = 0; }
        const uint32_t k = (uint32_t)st.seat;
        if (k & 1u) { lo0 |= h0; lo1 |= h1; }
        if (k & 2u) { hi0 |= h0; hi1 |= h1; }
        f0 &= ~h0; f1 &= ~h1;
    }
    pl.lo0 = lo0; pl.lo1 = lo1; pl.hi0 = hi0; pl.hi1 = hi1;
    if (q.n_rest) {                                           // deal-all 13 for the remaining seats
        uint32_t att = 0;
        bool need = valid && !bad;
        while (need) {
            const bool voided = deal_fixed_core(a.key, STREAM_REFORDER_REST + (att << 8), q.n_rest, q.q_rest, lo0, lo1, hi0, hi1,
                                                f0, f1, q.thresh_rest, il, ih, pl);
            if (!voided) need = false;
            else if (++att >= 64u) { bad = true; need = false; }
        }
    }
    return bad;
}

// Sampling kernel: a warp is a small dataflow machine.  Stage 1 runs on full warps of new indices and feeds a queue;
// each lane then carries ONE deal through the later ranged seats -- one rejection attempt per iteration, the narrowed
// ranges recomputed when it moves to the next seat -- and takes a new deal from the queue the moment its own is
// complete, so no lane waits for the slowest one of its warp; complete deals collect in a second queue and go through
// deal-all, tallies and the store 32 at a time.
constexpr int RQ_CAP = 64;
__global__ void __launch_bounds__(THREADS) sample_reforder_kernel(const __grid_constant__ SampleArgs a, const __grid_constant__ SeqDev q,
                                                                  const __grid_constant__ SeqRoot root)
{
    if (wave_cancelled(a)) return;
    __shared__ __align__(16) LowTabs s_low;
    __shared__ uint4 s_q1[WARPS_PER_BLOCK][RQ_CAP];           // stage-1 results: (hand m0, m1, index - a.first, -)
    __shared__ uint4 s_qd[WARPS_PER_BLOCK][RQ_CAP];           // complete deals: planes ...
    __shared__ uint4 s_qe[WARPS_PER_BLOCK][RQ_CAP];           // ... and (free f0, f1, index - a.first, -)
    const SmemPlan sp = smem_plan(false, false, false, false, a.tally_mask);
    uint32_t *smem = s_dyn;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    tally_init(a, sp, smem, FULL_HIST_COPIES);
    for (int i = threadIdx.x; i < (int)(sizeof(LowTabs) / 4); i += THREADS)
        reinterpret_cast<uint32_t *>(&s_low)[i] = reinterpret_cast<const uint32_t *>(&c_lowtabs)[i];
    __syncthreads();
    const TallyMem tm = tally_mem(sp, smem, warp);
    uint4 *q1 = s_q1[warp], *qd = s_qd[warp], *qe = s_qe[warp];
    unsigned h1 = 0, t1 = 0, hd = 0, td = 0;                  // queue cursors, warp-uniform
    const uint32_t lt = (1u << lane) - 1u;
    const unsigned long long total_warps = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    uint32_t n_raw = 0, n_bad = 0;
    // the deal this lane carries
    bool busy = false, fresh = false;
    uint32_t off = 0, lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0, f0 = 0, f1 = 0, j = 0, att = 0;
    SeqPacked pk = {0, 0, 0, 0, 0, 0};
    unsigned long long M1 = 0, M2 = 0, M4 = 0;                // honour weight planes of the remaining deck, compact coordinates
    int hcp_lo = 0, hcp_hi = 64;
    uint32_t batch = 0;

    auto run_done = [&](unsigned count) {                     // deal-all + tallies + store for up to 32 complete deals
        const uint4 d = qd[(hd + lane) % RQ_CAP], e = qe[(hd + lane) % RQ_CAP];
        hd += count;
        __syncwarp();
        const bool active = lane < count;
        const unsigned long long idx = a.first + e.z;
        Planes pl = {d.x, d.y, d.z, d.w};
        bool bad = false;
        if (q.n_rest) {
            uint32_t at = 0;
            bool need = active;
            while (need) {
                const bool voided = deal_fixed_core(a.key, STREAM_REFORDER_REST + (at << 8), q.n_rest, q.q_rest, d.x, d.y, d.z, d.w,
                                                    e.x, e.y, q.thresh_rest, (uint32_t)idx, (uint32_t)(idx >> 32), pl);
                if (!voided) need = false;
                else if (++at >= 64u) { bad = true; need = false; }
            }
        }
        n_bad += (active && bad) ? 1u : 0u;
        stage_b(a, pl, e.z, active && !bad, tm, lane, nullptr, 0u);      // nothing left to test
    };

    for (;;) {
        // 1. keep the stage-1 queue stocked: full warps of new indices
        while (t1 - h1 < 32u && batch < a.iters) {
            const unsigned long long o = ((unsigned long long)batch * total_warps + gwarp) * 32ull + lane;
            ++batch;
            const bool valid = o < a.n;
            const unsigned long long idx = a.first + o;
            uint32_t m0, m1;
            const bool ok = seq_stage1(a.key, root, s_low, (uint32_t)idx, (uint32_t)(idx >> 32), m0, m1) && valid;
            n_raw += valid ? 1u : 0u;
            n_bad += (valid && !ok) ? 1u : 0u;
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
            if (ok) q1[(t1 + __popc(bal & lt)) % RQ_CAP] = make_uint4(m0, m1, (uint32_t)o, 0u);
            t1 += __popc(bal);
            __syncwarp();
        }
        // 2. idle lanes take a deal from the queue
        {
            const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
            const unsigned avail = t1 - h1, want = __popc(idle);
            const unsigned take = want < avail ? want : avail;
            const unsigned my = __popc(idle & lt);
            if (!busy && my < take) {
                const uint4 e = q1[(h1 + my) % RQ_CAP];
                const uint32_t k = (uint32_t)q.st[0].seat;
                lo0 = a.fix.lo0 | ((k & 1u) ? e.x : 0u); lo1 = a.fix.lo1 | ((k & 1u) ? e.y : 0u);
                hi0 = a.fix.hi0 | ((k & 2u) ? e.x : 0u); hi1 = a.fix.hi1 | ((k & 2u) ? e.y : 0u);
                f0 = a.fix.free0 & ~e.x; f1 = a.fix.free1 & ~e.y;
                off = e.z; j = 1u; busy = true; fresh = true;
            }
            h1 += take;
            __syncwarp();
        }
        if (!__any_sync(0xFFFFFFFFu, busy)) {
            if (batch >= a.iters && t1 == h1) break;            // nothing in flight, nothing left to draw
            continue;
        }
        // 3. lanes that just received a deal or just finished a seat: done, or the next seat's narrowed ranges
        if (__any_sync(0xFFFFFFFFu, fresh)) {
            bool done = false;
            if (fresh) {
                if (j >= q.n_stages || q.st[j < 4u ? j : 3u].m <= 13u) done = true;      // a later seat is implied once 13 cards remain
                else {
                    const SeqRanges r = seq_infer(q.st[j], q.exhaust, f0, f1);
                    if (r.bad) { busy = false; ++n_bad; }      // the reference signals bad-range for this build
                    else {
                        pk = seq_pack(r); hcp_lo = r.hcp_lo; hcp_hi = r.hcp_hi; att = 0u;
                        seq_hcp_planes(f0, f1, M1, M2, M4);
                    }
                }
                fresh = false;
            }
            // complete deals queue up for deal-all (fewer than 32 wait there, at most 32 join: the queue holds 64)
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, done);
            if (bal) {
                if (done) {
                    const unsigned pos = (td + __popc(bal & lt)) % RQ_CAP;
                    qd[pos] = make_uint4(lo0, lo1, hi0, hi1);
                    qe[pos] = make_uint4(f0, f1, off, 0u);
                    busy = false;
                }
                td += __popc(bal);
                __syncwarp();
                while (td - hd >= 32u) run_done(32u);
            }
        }
        // 4. rejection: every busy lane draws SEQ_DRAWS candidates (cheap: one Philox call, Floyd's thirteen draws, HCP in
        //    compact coordinates) and keeps the first whose HCP fits; that one is deposited and fully tested.  Attempts
        //    are consumed strictly in order, so the accepted hand is the first admissible one of the index's stream --
        //    exactly what the one-thread form (reforder_deal) finds.
        if (__any_sync(0xFFFFFFFFu, busy)) {
            const SeqStage &st = q.st[j < 4u ? j : 3u];
            const unsigned long long idx = a.first + off;
            unsigned long long cand = 0ull;
            uint32_t cand_att = 0u;
            bool have = false;
            if (busy) {
#pragma unroll 1
                for (uint32_t d = 0; d < SEQ_DRAWS; ++d) {
                    unsigned long long c;
                    const bool clean = seq_draw_floyd(a.key, (uint32_t)idx, (uint32_t)(idx >> 32), STREAM_REFORDER + j, att + d, st.m, st.thresh, c);
                    const int hcp = seq_compact_hcp(c, M1, M2, M4);
                    if (!have && clean && hcp >= hcp_lo && hcp <= hcp_hi) { have = true; cand = c; cand_att = att + d; }
                }
            }
            bool ok = false;
            uint32_t x0 = 0, x1 = 0;
            if (have) {
                seq_expand(cand, f0, f1, x0, x1);
                ok = seq_packed_ok(pk, x0, x1);
            }
            if (busy) {
                if (ok) {
                    const uint32_t k = (uint32_t)st.seat;
                    if (k & 1u) { lo0 |= x0; lo1 |= x1; }
                    if (k & 2u) { hi0 |= x0; hi1 |= x1; }
                    f0 &= ~x0; f1 &= ~x1;
                    ++j;
                    fresh = true;
                } else {
                    att = have ? cand_att + 1u : att + SEQ_DRAWS;
                    if (att >= SEQ_MAX_ATTEMPTS) { busy = false; ++n_bad; }
                }
            }
        }
    }
    if (td - hd) run_done(td - hd);

    {
        unsigned long long r64 = n_raw, v64 = n_bad;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            r64 += __shfl_xor_sync(0xFFFFFFFFu, r64, o);
            v64 += __shfl_xor_sync(0xFFFFFFFFu, v64, o);
        }
        if (lane == 0) {
            atomicAdd(&a.tallies[0], r64);
            if (v64) atomicAdd(&a.tallies[1], v64);
        }
    }
    __syncthreads();
    flush_hist(a, sp, smem, FULL_HIST_COPIES);
}

// ---------------------------------------------------------------------------
// Expansion of index records (bmc_expand): the deal behind an index is a pure function of (seed, index,
// constraint, dealing mode), so an accepted deal can be shipped as its 32-bit index offset and regenerated on
// demand.  One thread per offset; mode 1 full deal (free or with fixed hands), 2 / 4 seat-first (count classes / exact
// class table), 3 reference order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) expand_kernel(const __grid_constant__ SampleArgs a, const __grid_constant__ SeqDev q,
                                                         const __grid_constant__ SeqRoot root, int mode,
                                                         const uint32_t *__restrict__ offsets, unsigned long long n, uint4 *__restrict__ out)
{
    __shared__ __align__(16) LowTabs s_low;
    if (mode >= 2) {
        for (int i = threadIdx.x; i < (int)(sizeof(LowTabs) / 4); i += THREADS)
            reinterpret_cast<uint32_t *>(&s_low)[i] = reinterpret_cast<const uint32_t *>(&c_lowtabs)[i];
        __syncthreads();
    }
    const unsigned long long stride = (unsigned long long)gridDim.x * THREADS;
    for (unsigned long long i = (unsigned long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
        const unsigned long long idx = a.first + offsets[i];
        Planes pl;
        if (mode == 3) reforder_deal(a, q, root, s_low, idx, true, pl);
        else if (mode == 2 || mode == 4) {
            uint32_t m0, m1;
            if (mode == 4) sf_exact_hand(a.key, a.cons.x, s_low, idx, true, m0, m1);
            else sf_primary_hand(a.key, a.cons.sf_tab, a.cons.sf_bmul, s_low, idx, true, m0, m1);
            const uint32_t P = a.cons.primary;
            const uint32_t xl = (P & 1u) ? 0xFFFFFFFFu : 0u, xh = (P & 2u) ? 0xFFFFFFFFu : 0u;
            deal_rest39(a.key, a.sf_q0, m0, m1, xl, xh, (uint32_t)idx, (uint32_t)(idx >> 32), pl);
        } else if (a.fix.n_free != 52u) deal_fixed(a.key, a.fix, (uint32_t)idx, (uint32_t)(idx >> 32), pl);
        else deal52(a.key, (uint32_t)idx, (uint32_t)(idx >> 32), pl);
        out[i] = make_uint4(pl.lo0, pl.lo1, pl.hi0, pl.hi1);
    }
}

// K2 (features / accept / choose-bid on given records) lives in filter.cuh: it is also compiled at run time with the
// rows of a rule table translated to C++ (NVRTC), like the sampling kernels.
#include "filter.cuh"
static_assert(sizeof(FeatOut) == sizeof(bmc_features), "FeatOut mirrors bmc_features");

// ---------------------------------------------------------------------------
// K2b: the auction driver of class `bidding` (`next`, bidding.cl:463-480) over a deal set: the first seat, from
// the dealer on, that has a bid over the opening scheme; the deal is "rolled" so that the opener sits first; the
// response scheme is the one the opening selects ((2 c) -> 2c-responses, NT -> (nt-responses level), a one-level
// suit -> (basic-responses bid), otherwise none) and the opener's partner chooses over it.
// ---------------------------------------------------------------------------
constexpr int AUCTION_MAX_OPENINGS = 32;
struct AuctionDev { SchemeDev open; SchemeDev resp[AUCTION_MAX_OPENINGS]; int dealer; };
struct AuctionOut { int8_t opener, opening, response, status; };        // same layout as bmc_auction_row
static_assert(sizeof(AuctionOut) == sizeof(bmc_auction_row), "AuctionOut mirrors bmc_auction_row");

// choose-bid over one table: first row whose form holds; -2 = an illegal form was reached first; -1 = nil
__device__ __forceinline__ int choose_row(const SchemeDev &sch, const uint8_t *fb)
{
    for (uint32_t row = 0; row < sch.n_rows; ++row) {
        bool err = false;
        const bool v = run_program(sch.code + __ldg(sch.row_off + row), fb, err);
        if (err) return -2;
        if (v) return (int)row;
    }
    return -1;
}

__global__ void __launch_bounds__(256) auction_kernel(const __grid_constant__ AuctionDev a, const uint4 *__restrict__ rec,
                                                      unsigned long long n, AuctionOut *__restrict__ out)
{
    __shared__ uint32_t s_feat[8 * FEAT_SCRATCH_WORDS];
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;                                   // the interpreter keeps per-lane state: lanes may diverge freely
    const uint4 r = rec[i];
    const Planes pl = {r.x, r.y, r.z, r.w};
    uint32_t *scratch = s_feat + (threadIdx.x >> 5) * FEAT_SCRATCH_WORDS;
    AuctionOut o = {-1, -1, -1, 0};
    for (int p = 0; p < 4 && o.opener < 0 && o.status == 0; ++p) {
        const int seat = (a.dealer + p) & 3;
        uint32_t m0, m1;
        seat_words(pl, seat, m0, m1);
        const int bid = choose_row(a.open, feat_store(scratch, threadIdx.x & 31u, hand_features(m0, m1)));
        if (bid == -2) o.status = 1;
        else if (bid >= 0) { o.opener = (int8_t)seat; o.opening = (int8_t)bid; }
    }
    if (o.opener >= 0) {
        const SchemeDev &rs = a.resp[o.opening];
        if (rs.n_rows) {
            uint32_t m0, m1;
            seat_words(pl, (o.opener + 2) & 3, m0, m1);
            const int bid = choose_row(rs, feat_store(scratch, threadIdx.x & 31u, hand_features(m0, m1)));
            if (bid == -2) o.status = 2;
            else o.response = (int8_t)bid;
        }
    }
    out[i] = o;
}

// ---------------------------------------------------------------------------
// K4: scoring
// ---------------------------------------------------------------------------
// contract-score / base-score / match-points and the stand-in trick source live in score.cuh (shared with the
// sampling kernels, which tally scores over the accepted deals in stage B).
static_assert(sizeof(ContractDev) == sizeof(bmc_contract), "ContractDev mirrors bmc_contract");

// HBM-bound (1 B in, 4 B out per element): each thread scores four consecutive elements
// (one 32-bit load, one 128-bit store); the contract index advances without a division.
__global__ void __launch_bounds__(256) score_kernel(const ContractDev *__restrict__ contracts, const uint8_t *__restrict__ vul,
                                                    uint32_t nc, const uint8_t *__restrict__ tricks, unsigned long long total,
                                                    int32_t *__restrict__ scores)
{
    __shared__ ContractDev s_c[64];
    __shared__ uint8_t s_v[64];
    const bool cached = nc <= 64u;
    if (cached)
        for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) { s_c[i] = contracts[i]; s_v[i] = vul ? vul[i] : 0; }
    __syncthreads();
    const unsigned long long quads = total >> 2, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
        const uint32_t t4 = reinterpret_cast<const uint32_t *>(tricks)[q];
        uint32_t c = (uint32_t)((q << 2) % nc);
        int4 out;
        int *o = &out.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const ContractDev ct = cached ? s_c[c] : contracts[c];
            const int v = cached ? s_v[c] : (vul ? vul[c] : 0);
            o[k] = base_score_dev(ct, (int)((t4 >> (8 * k)) & 0xFFu), v);
            c = c + 1u == nc ? 0u : c + 1u;
        }
        reinterpret_cast<int4 *>(scores)[q] = out;
    }
    // tail (total not a multiple of 4)
    for (unsigned long long i = (quads << 2) + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t c = (uint32_t)(i % nc);
        scores[i] = base_score_dev(contracts[c], tricks[i], vul ? vul[c] : 0);
    }
}

__global__ void __launch_bounds__(256) imps_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b,
                                                   unsigned long long n, int8_t *__restrict__ imps, uint8_t *__restrict__ status)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool bad;
        const int v = imps_dev(a[i] - b[i], bad);
        imps[i] = (int8_t)v;
        if (status) status[i] = bad ? 1 : 0;
    }
}

// Index-keyed drill tally: tricks of deal `idx` = tricks_of(Philox(seed; idx_lo, idx_hi, block, stream 1)) (score.cuh):
// eight exactly uniform values 0..13.  hist = tricks[8][14] | imps[8][49] | dropped[8] (u64 words).
__global__ void __launch_bounds__(256) score_tally_kernel(const __grid_constant__ ContractTable t, const __grid_constant__ PhiloxKey key,
                                                          unsigned long long first, unsigned long long n,
                                                          unsigned long long *__restrict__ hist)
{
    __shared__ uint32_t s_h[SCORE_WORDS];
    __shared__ int s_score[SCORE_TR_WORDS];
    for (int i = threadIdx.x; i < SCORE_TR_WORDS; i += blockDim.x) {
        const int c = i / 14;
        s_score[i] = c < (int)t.n ? base_score_dev(t.c[c], i % 14, t.vul[c]) : 0;
    }
    for (int i = threadIdx.x; i < SCORE_WORDS; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long idx = first + i;
        uint32_t tr[8];
        tricks_of(key, (uint32_t)idx, (uint32_t)(idx >> 32), 0u, 1u, tr);
        score_tally_update(t, s_score, tr, s_h);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SCORE_WORDS; i += blockDim.x)
        if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

// ---------------------------------------------------------------------------
// INT32 issue-rate microbenchmark: per thread, 8 independent chains mixing the
// fma pipe (IMAD) and the alu pipe (LOP3 / IADD3), the two pipes the deal uses.
// ---------------------------------------------------------------------------
constexpr int PEAK_ITERS = 1024, PEAK_OPS_PER_ITER = 128;
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t seed, uint32_t *out)
{
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u;
    uint32_t b0 = a0 ^ 0x1234u, b1 = a1 ^ 0x5678u, b2 = a2 ^ 0x9ABCu, b3 = a3 ^ 0xDEF0u;
#pragma unroll 1
    for (int i = 0; i < PEAK_ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            a0 = a0 * 0x010101u + b0; b0 = b0 ^ a1 ^ 0x9E3779B9u;
            a1 = a1 * 0x010103u + b1; b1 = b1 ^ a2 ^ 0x7F4A7C15u;
            a2 = a2 * 0x010105u + b2; b2 = b2 ^ a3 ^ 0x85EBCA6Bu;
            a3 = a3 * 0x010107u + b3; b3 = b3 ^ a0 ^ 0xC2B2AE35u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
}

// Runs after every wave.  Publishes the four leading counters and a wave sequence number into host-mapped
// pinned memory -- so the host follows the job without a device->host copy (the copy engine streams the
// records) and without synchronising the stream -- and snapshots the accepted counter for the next wave's
// cancellation test (wave_cancelled, kernels.cuh).
__global__ void publish_counters_kernel(const unsigned long long *__restrict__ tallies, volatile unsigned long long *mapped,
                                        unsigned long long *snap, unsigned long long seq, const unsigned long long *extra)
{
    if (threadIdx.x < 4) mapped[threadIdx.x] = tallies[threadIdx.x];
    if (threadIdx.x == 4) mapped[5] = extra ? *extra : 0ull;          // byte cursor of the packed index records
    if (threadIdx.x == 0) *snap = tallies[2];
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) mapped[4] = seq;
    __threadfence_system();
}

// Sums the tally buffers of peer devices into this device's buffer over NVLink (bmc_sample_multi): P2P loads.
struct PeerTallies { const unsigned long long *p[16]; int n; };
__global__ void reduce_peer_tallies_kernel(unsigned long long *dst, const __grid_constant__ PeerTallies peers, int words)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) {
        unsigned long long v = dst[i];
        for (int g = 0; g < peers.n; ++g) v += peers.p[g][i];
        dst[i] = v;
    }
}

// ---------------------------------------------------------------------------
// host objects
// ---------------------------------------------------------------------------
struct bmc_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    unsigned long long *h_head = nullptr;        // pinned + mapped: {raw, voided, accepted, stored, wave sequence number}
    unsigned long long *h_head_dev = nullptr;    // its device alias
    unsigned long long *d_snap = nullptr;        // accepted counter at the end of the previous wave
    void *d_rec = nullptr;                       // grow-only record scratch of the host-buffer API
    uint64_t rec_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long *d_tallies = nullptr;     // scratch tallies for the host-buffer API
    uint8_t *d_arena = nullptr;                  // grow-only per-deal arenas of the trick estimator (bmc_play_deal)
    uint64_t arena_bytes = 0;
    uint32_t *d_bitmap = nullptr;                // packed index records (bmc_sample_packed): one wave's bitmap,
    uint64_t bitmap_bytes = 0;
    uint8_t *d_packed = nullptr;                 // the byte stream,
    uint64_t packed_bytes = 0;
    PackState *d_pack = nullptr;                 // and the cursors
    uint64_t launches = 0;
    std::mutex mu;
};

struct ConsTables {              // per-device copies of a constraint's tables
    uint32_t *d_code = nullptr;
    JitKernel jit[3];            // [0] full-deal kernel, [1] seat-first kernel, [2] seat-first kernel, exact class table
    // exact class table of the primary seat (mode 4; distgen.cuh)
    unsigned long long *x_cum = nullptr;
    uint32_t *x_keys = nullptr, *x_bkt = nullptr;
    uint32_t x_bmul = 0, x_nbkt = 0;
    bool x_built = false;
    SeqRoot root{};              // mode 3: the root generator's class table (uses the x_* arrays) and its scale
};

struct bmc_constraints {
    ConsDev dev{};               // dev.code filled per device
    std::vector<uint32_t> code;
    std::map<int, ConsTables> tabs;      // device -> tables (guarded by mu)
    uint32_t tab_off = 0;        // seat-first class table: word offset in `code`
    int sf_hcp[2] = {0, 37};     // the HCP window the table was built for (primary seat)
    bool has_cons = false;
    bool has_fixed = false;
    FixDev fix{};
    bmc_seat_spec specs[4];      // as given (the reference-order sampler re-reads them)
    bool has_rules = false;
    int ref_defs = -1;           // reference-order mode: number of ranged seats after the fixed ones
    SeqDev seq{};
    std::string rule_text[4];    // the rule forms as given (the NVRTC path re-reads them)
    int jit = 0;                 // 0: interpret the rule bytecode; 1: compile the rule forms into the kernel (NVRTC); 2: auto
    float jit_ms = 0.f;
    double interp_ms = 0.0;      // kernel time spent so far with interpreted rules (auto-JIT trigger)
    std::string jit_log;
    int mode = 0;                // requested: 0 auto, 1 full deal (deal52), 2 seat-first, 3 reference order
    int resolved = 0;            // 0 = not calibrated yet
    double stage_a_pass = -1.0;  // pass rate of the primary seat (exact: x_weight / C(52,13))
    uint64_t x_weight = 0;       // hands the primary seat's constraint admits (exact, from the device enumeration)
    uint64_t x_classes = 0;      // classes (honours x low-card counts) among them
    bool x_known = false;
    ContractTable score{};       // bmc_constraints_set_scoring
    std::mutex mu, cal_mu;
};

struct SchemeTables { uint32_t *d_code = nullptr, *d_off = nullptr; JitKernel jit; };
struct bmc_scheme {
    std::vector<uint32_t> code, row_off;
    std::vector<std::string> bids;
    std::vector<Sx> forms;       // the rows' rule forms (the NVRTC path re-reads them)
    std::map<int, SchemeTables> tabs;
    int jit = 2;                 // 0 interpret, 1 compile (NVRTC), 2 auto: compile for large or repeated jobs
    double interp_ms = 0.0;      // kernel time spent interpreting this table
    float jit_ms = 0.f;
    std::mutex mu;
};

static int check_contracts(const bmc_contract *c, uint32_t n)
{
    for (uint32_t i = 0; i < n; ++i)
        if (c[i].level < 0 || c[i].level > 7 || c[i].suit < 0 || c[i].suit > 4 || c[i].dbl < 0 || c[i].dbl > 2 || c[i].declarer < -1 || c[i].declarer > 3)
            return fail(BMC_E_INVALID, "contract " + std::to_string(i) + " out of range");
    return BMC_OK;
}

static int fill_contract_table(ContractTable &t, const bmc_contract *contracts, const uint8_t *vulnerable, uint32_t nc,
                               const uint8_t *pairs, uint32_t n_pairs)
{
    { int rc = check_contracts(contracts, nc); if (rc) return rc; }
    t.n = nc; t.n_pairs = n_pairs;
    for (uint32_t i = 0; i < nc; ++i) {
        t.c[i].level = contracts[i].level; t.c[i].suit = contracts[i].suit;
        t.c[i].declarer = contracts[i].declarer; t.c[i].dbl = contracts[i].dbl;
        t.vul[i] = vulnerable ? (vulnerable[i] ? 1 : 0) : 0;
    }
    for (uint32_t i = 0; i < 2 * n_pairs; ++i) {
        if (pairs[i] >= nc) return fail(BMC_E_INVALID, "pair index out of range");
        t.pairs[i] = pairs[i];
    }
    return BMC_OK;
}

static const ShapeList g_shapes = make_shape_list();

extern "C" {

const char *bmc_last_error(void) { return g_err.c_str(); }
const char *bmc_version(void) { return "bridgemc 0.2 (sm_100a)"; }

void bmc_shape_table(uint8_t out[39][4]) { memcpy(out, g_shapes.s, sizeof g_shapes.s); }

static int ctx_setup(bmc_ctx *c)
{
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
    c->sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaMalloc(&c->d_tallies, sizeof(bmc_tallies)));
    CUDA_TRY(cudaMalloc(&c->d_snap, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(c->d_snap, 0, sizeof(unsigned long long)));
    CUDA_TRY(cudaHostAlloc(&c->h_head, 8 * sizeof(unsigned long long), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer((void **)&c->h_head_dev, c->h_head, 0));
    return BMC_OK;
}

int bmc_init(int device, bmc_ctx **out)
{
    if (!out) return fail(BMC_E_INVALID, "bmc_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(BMC_E_CUDA, std::string("no CUDA device (") + cudaGetErrorString(e) + "); libbridgemc has no CPU fallback");
    if (device < 0 || device >= count) return fail(BMC_E_INVALID, "bmc_init: device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    bmc_ctx *c = new bmc_ctx();
    c->device = device;
    const int rc = ctx_setup(c);
    if (rc != BMC_OK) {                     // a failed init releases whatever it had created (g_err keeps the cause)
        const std::string why = g_err;
        bmc_shutdown(c);
        return fail(rc, why);
    }
    *out = c;
    return BMC_OK;
}

void bmc_shutdown(bmc_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->d_tallies) cudaFree(c->d_tallies);
    if (c->d_arena) cudaFree(c->d_arena);
    if (c->d_bitmap) cudaFree(c->d_bitmap);
    if (c->d_packed) cudaFree(c->d_packed);
    if (c->d_pack) cudaFree(c->d_pack);
    if (c->d_snap) cudaFree(c->d_snap);
    if (c->h_head) cudaFreeHost(c->h_head);
    if (c->d_rec) cudaFree(c->d_rec);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int bmc_set_stream(bmc_ctx *c, void *s)
{
    if (!c) return fail(BMC_E_INVALID, "ctx is NULL");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return BMC_OK;
}

uint64_t bmc_launch_count(const bmc_ctx *c) { return c ? c->launches : 0; }

void *bmc_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
void bmc_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- constraints ------------------------------------------------------------
static int spec_to_ranges(const bmc_seat_spec &sp, SeatRanges &r, std::string &why)
{
    auto apply = [&](int f, int lo, int hi, int cap) {
        if (lo >= 0 && lo > r.lo[f]) r.lo[f] = lo;
        if (hi >= 0 && hi < r.hi[f]) r.hi[f] = hi;
        (void)cap;
    };
    apply(F_HCP, sp.hcp[0], sp.hcp[1], 37);
    for (int s = 0; s < 4; ++s) {
        apply(F_C + s, sp.len[s][0], sp.len[s][1], 13);
        apply(F_CP + s, sp.suit_hcp[s][0], sp.suit_hcp[s][1], 10);
    }
    apply(F_LOS, sp.losers[0], sp.losers[1], 12);
    for (int f = 0; f < F_BAL; ++f)
        if (r.lo[f] > r.hi[f]) {
            why = "empty range for feature " + std::to_string(f) + ": <" + std::to_string(r.lo[f]) + "-" + std::to_string(r.hi[f]) + ">";
            return BMC_E_BAD_RANGE;
        }
    return BMC_OK;
}

static void ranges_to_seatc(const SeatRanges &r, SeatC &c)
{
    int lo[12], hi[12];
    for (int f = 0; f < 12; ++f) { lo[f] = 0; hi[f] = 127; }
    for (int f = 0; f < F_BAL; ++f) { lo[f] = r.lo[f]; hi[f] = r.hi[f] > 127 ? 127 : r.hi[f]; }
    lo[F_BAL] = r.bal_req == 1 ? 1 : 0;
    hi[F_BAL] = r.bal_req == 2 ? 0 : 1;
    c.need = 0;
    for (int s = 0; s < 4; ++s) if (r.lo[F_CP + s] > 0 || r.hi[F_CP + s] < 10) c.need |= 2u;
    c.hcp_lo_span = (uint32_t)r.lo[F_HCP] | ((uint32_t)(hi[F_HCP] - r.lo[F_HCP]) << 8);
    for (int w = 0; w < 3; ++w) {
        uint32_t A = 0, B = 0;
        for (int b = 0; b < 4; ++b) {
            const int f = 4 * w + b;
            A |= (uint32_t)(0x80 - lo[f]) << (8 * b);
            B |= (uint32_t)(0x7F - hi[f]) << (8 * b);
        }
        c.A[w] = A;
        c.B[w] = B;
    }
}

// Class table of the seat-first stage A1 (seatfirst.cuh): the 625 honour-count classes (nJ nQ nK nA) of a
// 13-card hand, those with hcp_lo <= HCP <= hcp_hi first, each with S * (weight before it), its counts,
// I = product of C(4, n.) and floor(2^32 / I).  Appended to the code buffer (16-byte aligned).
static void build_rank_table(bmc_constraints *c, int hcp_lo, int hcp_hi)
{
    static const uint64_t C4[5] = {1, 4, 6, 4, 1};
    uint64_t C36[14];
    C36[0] = 1;
    for (int m = 1; m <= 13; ++m) C36[m] = C36[m - 1] * (uint64_t)(36 - m + 1) / (uint64_t)m;
    const uint64_t N = 635013559600ull;                               // C(52,13)
    const uint64_t S = (uint64_t)((((unsigned __int128)1) << 64) / N);
    struct Cls { uint32_t info, magic; uint64_t w; };
    std::vector<Cls> acc, rej;
    for (int id = 0; id < 625; ++id) {
        const int nJ = id % 5, nQ = (id / 5) % 5, nK = (id / 25) % 5, nA = id / 125, n = nJ + nQ + nK + nA;
        if (n > 13) continue;
        const uint64_t I = C4[nJ] * C4[nQ] * C4[nK] * C4[nA];
        const int hcp = nJ + 2 * nQ + 3 * nK + 4 * nA;
        Cls k;
        k.info = (uint32_t)nJ | ((uint32_t)nQ << 3) | ((uint32_t)nK << 6) | ((uint32_t)nA << 9) | ((uint32_t)I << 12);
        k.magic = I == 1 ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / I);
        k.w = I * C36[13 - n];
        (hcp >= hcp_lo && hcp <= hcp_hi ? acc : rej).push_back(k);
    }
    while (c->code.size() & 3u) c->code.push_back(enc(OP_END));
    c->tab_off = (uint32_t)c->code.size();
    uint64_t cum = 0;
    int entries = 0;
    auto emit = [&](const Cls &k) {
        const uint64_t v = cum * S;
        c->code.push_back((uint32_t)v); c->code.push_back((uint32_t)(v >> 32));
        c->code.push_back(k.info); c->code.push_back(k.magic);
        cum += k.w;
        ++entries;
    };
    for (const Cls &k : acc) emit(k);
    const uint64_t w_acc = cum;
    for (const Cls &k : rej) emit(k);
    for (; entries < RANK_TAB_ENTRIES; ++entries) {
        c->code.push_back(0xFFFFFFFFu); c->code.push_back(0xFFFFFFFFu); c->code.push_back(1u << 12); c->code.push_back(0xFFFFFFFFu);
    }
    // bucket index of the class search (seatfirst.cuh sf_find_class)
    const uint64_t theta = w_acc * S;
    const uint64_t bm = ((uint64_t)SF_BUCKETS << 32) / ((theta >> 32) + 1);
    const uint32_t bmul = bm > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)bm;
    const size_t tab_words = (size_t)c->tab_off;
    auto cum_of = [&](int k) {
        return (uint64_t)c->code[tab_words + 4 * (size_t)k] | ((uint64_t)c->code[tab_words + 4 * (size_t)k + 1] << 32);
    };
    auto f = [&](uint64_t u_hi) { return (uint32_t)((u_hi * (uint64_t)bmul) >> 32); };
    int pos = 0;
    for (uint32_t j = 0; j < (uint32_t)SF_BUCKETS; j += 2) {
        uint32_t pair16[2];
        for (uint32_t h = 0; h < 2; ++h) {
            // smallest u_hi with f(u_hi) >= j + h
            uint64_t lo = (((uint64_t)(j + h) << 32) + bmul - 1) / bmul;
            while (lo > 0 && f(lo - 1) >= j + h) --lo;
            while (lo <= 0xFFFFFFFFull && f(lo) < j + h) ++lo;
            const uint64_t u_min = lo > 0xFFFFFFFFull ? ~0ull : lo << 32;
            while (pos + 1 < RANK_TAB_ENTRIES - 1 && cum_of(pos + 1) <= u_min) ++pos;
            pair16[h] = (uint32_t)pos;
        }
        c->code.push_back(pair16[0] | (pair16[1] << 16));
    }
    c->dev.sf_bmul = bmul;
    c->dev.sf_theta = w_acc * S;
    c->dev.sf_void = N * S;
    c->sf_hcp[0] = hcp_lo; c->sf_hcp[1] = hcp_hi;
}

int bmc_constraints_create(const bmc_seat_spec specs[4], const char *const rules[4], bmc_constraints **out)
{
    if (!out) return fail(BMC_E_INVALID, "out is NULL");
    *out = nullptr;
    std::unique_ptr<bmc_constraints> c(new bmc_constraints());
    RuleCompiler rc;
    uint64_t fixed_union = 0;
    uint64_t fixed_mask[4] = {0, 0, 0, 0};
    bool is_fixed[4] = {false, false, false, false};
    int total_lo[F_COUNT] = {0};
    double selectivity[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) {
        memset(&c->specs[k], 0xFF, sizeof(bmc_seat_spec));        // every bound -1 (nil)
        c->specs[k].fixed = 0;
        c->specs[k].fixed_mask = 0;
        if (specs) c->specs[k] = specs[k];
        if (rules && rules[k] && rules[k][0]) { c->has_rules = true; c->rule_text[k] = rules[k]; }
    }
    for (int k = 0; k < 4; ++k) {
        SeatC &sc = c->dev.seat[k];
        sc.prog = 0xFFFFFFFFu;
        sc.active = 0;
        SeatRanges r;
        bool has_spec = false;
        if (specs) {
            const bmc_seat_spec &sp = specs[k];
            if (sp.fixed) {
                if ((sp.fixed_mask & ~BMC_LANE_MASK) || __builtin_popcountll(sp.fixed_mask) != 13)
                    return fail(BMC_E_INVALID, "fixed hand of seat " + std::to_string(k) + " is not 13 cards in lane-mask layout");
                if (sp.fixed_mask & fixed_union) return fail(BMC_E_INVALID, "fixed hands overlap");
                fixed_union |= sp.fixed_mask;
                fixed_mask[k] = sp.fixed_mask;
                is_fixed[k] = true;
                c->has_fixed = true;
            }
            std::string why;
            int rcode = spec_to_ranges(sp, r, why);
            if (rcode != BMC_OK) return fail(rcode, "seat " + std::to_string(k) + ": " + why);
            for (int f = 0; f < F_BAL; ++f) if (r.lo[f] > 0 || r.hi[f] < 63) has_spec = true;
        }
        bool has_rule = rules && rules[k] && rules[k][0];
        bool exact = true;
        if (has_rule) {
            try {
                SxReader rd(rules[k]);
                Sx form = rd.read();
                if (form.is_list() && form.items.size() == 2 && form.items[0].is_sym("QUOTE")) form = form.items[1];
                std::vector<uint32_t> code;
                rc.compile(form, code, &r, &exact);
                if (!exact) {
                    sc.prog = (uint32_t)c->code.size();
                    c->code.insert(c->code.end(), code.begin(), code.end());
                }
            } catch (const RuleError &e) {
                return fail(BMC_E_PARSE, "seat " + std::to_string(k) + " rule: " + e.what());
            }
            if (r.empty())
                return fail(BMC_E_BAD_RANGE, "seat " + std::to_string(k) + ": rule and ranges exclude every hand");
        }
        // clamp to what a 13-card hand can hold (keeps the byte arithmetic in range)
        for (int s = 0; s < 4; ++s) { if (r.hi[F_C + s] > 13) r.hi[F_C + s] = 13; if (r.hi[F_CP + s] > 10) r.hi[F_CP + s] = 10; }
        if (r.hi[F_HCP] > 37) r.hi[F_HCP] = 37;
        if (r.hi[F_LOS] > 12) r.hi[F_LOS] = 12;
        if (r.empty()) return fail(BMC_E_BAD_RANGE, "seat " + std::to_string(k) + ": empty range");
        ranges_to_seatc(r, sc);
        sc.active = (has_spec || has_rule) ? 1u : 0u;
        if (sc.active) {
            c->has_cons = true;
            c->dev.n_active++;
            // crude selectivity score for choosing the stage-A seat: narrower hcp and
            // length windows reject more
            double sel = (double)(r.hi[F_HCP] - r.lo[F_HCP] + 1) / 38.0;
            for (int s = 0; s < 4; ++s) sel *= (double)(r.hi[F_C + s] - r.lo[F_C + s] + 1) / 14.0;
            if (r.bal_req) sel *= 0.5;
            selectivity[k] = sel;
            for (int f = 0; f < F_BAL; ++f) total_lo[f] += r.lo[f];
        }
    }
    // necessary joint conditions (the sums the reference's infer-* narrow with,
    // bridge.cl:1126-1234): minimum HCP over all seats <= 40, minimum lengths <= 13
    if (total_lo[F_HCP] > 40) return fail(BMC_E_BAD_RANGE, "Can't infer correct limit!: minimum HCP of the seats exceeds 40");
    for (int s = 0; s < 4; ++s)
        if (total_lo[F_C + s] > 13) return fail(BMC_E_BAD_RANGE, "Can't infer correct limit!: minimum suit lengths exceed 13");
    c->dev.primary = 0;
    c->dev.n_prog = 0;
    c->dev.b1_mask = 0xFu;
    c->dev.pad = 0;
    for (int k = 0; k < 4; ++k)
        if (c->dev.seat[k].active && c->dev.seat[k].prog != 0xFFFFFFFFu) c->dev.n_prog++;
    double best = 2.0;
    for (int k = 0; k < 4; ++k)
        if (c->dev.seat[k].active && selectivity[k] < best) { best = selectivity[k]; c->dev.primary = (uint32_t)k; }
    c->code.push_back(enc(OP_END));
    if (c->has_cons) {
        const uint32_t ls = c->dev.seat[c->dev.primary].hcp_lo_span;
        build_rank_table(c.get(), (int)(ls & 0xFFu), (int)((ls & 0xFFu) + (ls >> 8)));
    }
    if (c->has_fixed && is_fixed[0] && is_fixed[1] && is_fixed[2] && is_fixed[3])
        return fail(BMC_E_INVALID, "all four hands are fixed: nothing to deal");
    fixdev_init(c->fix, fixed_mask, is_fixed);
    *out = c.release();
    return BMC_OK;
}

// A change of dealing mode invalidates the per-device class tables (mode 4's and mode 3's share the arrays).
static void drop_class_tables(bmc_constraints *c)
{
    std::lock_guard<std::mutex> lk(c->mu);
    for (auto &kv : c->tabs) {
        ConsTables &t = kv.second;
        if (t.x_cum || t.x_keys || t.x_bkt) {
            cudaSetDevice(kv.first);
            cudaFree(t.x_cum); cudaFree(t.x_keys); cudaFree(t.x_bkt);
        }
        t.x_cum = nullptr; t.x_keys = nullptr; t.x_bkt = nullptr;
        t.x_built = false;
        t.root = SeqRoot{};
    }
}

int bmc_constraints_set_mode(bmc_constraints *c, int mode)
{
    if (!c || mode < 0 || mode > 4 || mode == 3) return fail(BMC_E_INVALID, "bmc_constraints_set_mode: bad argument (mode 3 is set by bmc_constraints_set_reference_order)");
    if ((mode == 2 || mode == 4) && (c->has_fixed || !c->has_cons)) return fail(BMC_E_INVALID, "seat-first modes need a constrained seat and no fixed hands");
    drop_class_tables(c);
    c->mode = mode;
    c->resolved = 0;
    return BMC_OK;
}
// Builds the per-stage constants of the reference-order sampler from the stored specs.
static int build_seq(bmc_constraints *c, int n_defs)
{
    int n_fixed = 0;
    while (n_fixed < 4 && c->specs[n_fixed].fixed) ++n_fixed;
    for (int k = n_fixed; k < 4; ++k)
        if (c->specs[k].fixed) return fail(BMC_E_INVALID, "reference order: fixed (:=) hands must occupy the first seats, as in `build`");
    if (n_defs < 1 || n_fixed + n_defs > 4) return fail(BMC_E_INVALID, "reference order: bad number of ranged seats");
    if (c->has_rules) return fail(BMC_E_INVALID, "reference order: the reference generator takes ranges only, not rule forms");
    SeqDev &q = c->seq;
    memset(&q, 0, sizeof q);
    q.exhaust = (n_fixed + n_defs == 4) ? 1u : 0u;
    auto has_hcp = [](const bmc_seat_spec &sp) { return sp.hcp[0] >= 0 || sp.hcp[1] >= 0; };
    auto has_suit = [](const bmc_seat_spec &sp, int s) {
        return sp.len[s][0] >= 0 || sp.len[s][1] >= 0 || sp.suit_hcp[s][0] >= 0 || sp.suit_hcp[s][1] >= 0;
    };
    auto has_third = [](const bmc_seat_spec &sp, int s) { return sp.suit_hcp[s][0] >= 0 || sp.suit_hcp[s][1] >= 0; };
    uint32_t m = 52u - 13u * (uint32_t)n_fixed;
    uint32_t cap[4] = {13, 13, 13, 13};
    for (int k = 0; k < n_fixed; ++k) cap[k] = 0;
    int sampled = 0;
    for (int j = 0; j < n_defs; ++j, m -= 13u) {
        SeqStage &st = q.st[j];
        const bmc_seat_spec &own = c->specs[n_fixed + j];
        st.seat = (int8_t)(n_fixed + j);
        st.m = m;
        st.has_hcp = has_hcp(own) ? 1 : 0;
        st.hcp_lo = own.hcp[0]; st.hcp_hi = own.hcp[1];
        st.has_others = j + 1 < n_defs ? 1 : 0;
        int su = 0, sl = 0;
        for (int o = j + 1; o < n_defs; ++o) {
            const bmc_seat_spec &ot = c->specs[n_fixed + o];
            if (!has_hcp(ot)) { st.others_missing_hcp = 1; continue; }
            if (ot.hcp[1] >= 0) su += ot.hcp[1];
            if (ot.hcp[0] >= 0) sl += ot.hcp[0];
        }
        st.sum_upper_hcp = (int16_t)su; st.sum_lower_hcp = (int16_t)sl;
        for (int s = 0; s < 4; ++s) {
            SeqSuit &qs = st.suit[s];
            qs.has_spec = has_suit(own, s) ? 1 : 0;
            qs.lo = own.len[s][0]; qs.hi = own.len[s][1];
            qs.has_third = has_third(own, s) ? 1 : 0;
            qs.third_lo = own.suit_hcp[s][0]; qs.third_hi = own.suit_hcp[s][1];
            int u = 0, l = 0, lt = 0;
            for (int o = j + 1; o < n_defs; ++o) {
                const bmc_seat_spec &ot = c->specs[n_fixed + o];
                if (!has_suit(ot, s)) { qs.others_missing = 1; continue; }
                // (or (second val) (first val)) for the lower-bound fold, (first val) for the upper-bound fold
                if (ot.len[s][1] >= 0) u += ot.len[s][1]; else if (ot.len[s][0] >= 0) u += ot.len[s][0];
                if (ot.len[s][0] >= 0) l += ot.len[s][0];
                if (has_third(ot, s) && ot.suit_hcp[s][0] >= 0) lt += ot.suit_hcp[s][0];
            }
            qs.sum_upper = (int16_t)u; qs.sum_lower = (int16_t)l; qs.sum_lower_third = (int16_t)lt;
        }
        // Floyd's 13 bounded draws (ranges m-12 .. m), four per Philox word, the last one alone (reforder.cuh seq_draw_floyd)
        for (int g = 0; g < 4; ++g) {
            uint64_t N = 1;
            for (int d = 0; d < (g < 3 ? 4 : 1); ++d) N *= (uint64_t)((int)m - 12 + 4 * g + d);
            st.thresh[g] = (uint32_t)((1ull << 32) % N);
        }
        for (int g = 4; g < 13; ++g) st.thresh[g] = 0;
        if (j == 0 || m > 13u) { cap[n_fixed + j] = 0; ++sampled; }
    }
    q.n_stages = (uint32_t)n_defs;
    q.n_rest = 52u - 13u * (uint32_t)(n_fixed + sampled);
    const uint32_t p0 = cap[0], p1 = p0 + cap[1], p2 = p1 + cap[2];
    q.q_rest = (128u - p0) | ((128u - p1) << 8) | ((128u - p2) << 16);
    for (int w = 0; w < 10; ++w) {
        uint64_t N = 1;
        for (int d = 0; d < 4; ++d) {
            const int n = (int)q.n_rest - 4 * w - d;
            if (n >= 2) N *= (uint64_t)n;
        }
        q.thresh_rest[w] = (uint32_t)((1ull << 32) % N);
    }
    c->ref_defs = n_defs;
    return BMC_OK;
}

int bmc_constraints_set_reference_order(bmc_constraints *c, int n_defs)
{
    if (!c) return fail(BMC_E_INVALID, "bmc_constraints_set_reference_order: NULL");
    int rc = build_seq(c, n_defs);
    if (rc) return rc;
    drop_class_tables(c);
    c->mode = 3;
    c->resolved = 3;
    return BMC_OK;
}
int bmc_constraints_mode(const bmc_constraints *c) { return c ? c->resolved : 0; }
int bmc_constraints_primary(const bmc_constraints *c) { return c ? (int)c->dev.primary : -1; }
int bmc_constraints_primary_hcp(const bmc_constraints *c, int32_t *lo, int32_t *hi)
{
    if (!c || !lo || !hi) return fail(BMC_E_INVALID, "bmc_constraints_primary_hcp: NULL argument");
    *lo = c->sf_hcp[0];
    *hi = c->sf_hcp[1];
    return BMC_OK;
}

void bmc_constraints_destroy(bmc_constraints *c)
{
    if (!c) return;
    for (auto &kv : c->tabs) {
        cudaSetDevice(kv.first);
        if (kv.second.d_code) cudaFree(kv.second.d_code);
        if (kv.second.x_cum) cudaFree(kv.second.x_cum);
        if (kv.second.x_keys) cudaFree(kv.second.x_keys);
        if (kv.second.x_bkt) cudaFree(kv.second.x_bkt);
        for (JitKernel &jk : kv.second.jit)
            if (jk.module) JitApi::get().ModuleUnload(jk.module);
    }
    delete c;
}

// The contracts / IMP pairs tallied over the accepted deals when tally_mask has BMC_TALLY_SCORE.
int bmc_constraints_set_scoring(bmc_constraints *c, const bmc_contract *contracts, const uint8_t *vulnerable, uint32_t nc,
                                const uint8_t *pairs, uint32_t n_pairs)
{
    if (!c || (nc && !contracts) || nc > (uint32_t)SCORE_MAX_CONTRACTS || n_pairs > (uint32_t)SCORE_MAX_PAIRS || (n_pairs && !pairs))
        return fail(BMC_E_INVALID, "bmc_constraints_set_scoring: bad argument");
    ContractTable t{};
    int rc = fill_contract_table(t, contracts, vulnerable, nc, pairs, n_pairs);
    if (rc) return rc;
    c->score = t;
    return BMC_OK;
}

// ---- NVRTC specialisation of a constraint (jit.hpp) ------------------------------------------------
static const char kJitPrelude[] =
    "typedef unsigned char uint8_t; typedef signed char int8_t; typedef unsigned short uint16_t; typedef short int16_t;\n"
    "typedef unsigned int uint32_t; typedef int int32_t; typedef unsigned long long uint64_t; typedef long long int64_t;\n"
    "#define BMC_JIT 1\n";
static const char kJitVars[] =
    "    const int vC = f.f0 & 0xFF, vD = (f.f0 >> 8) & 0xFF, vH = (f.f0 >> 16) & 0xFF, vS = (f.f0 >> 24) & 0xFF;\n"
    "    const int vCP = f.f1 & 0xFF, vDP = (f.f1 >> 8) & 0xFF, vHP = (f.f1 >> 16) & 0xFF, vSP = (f.f1 >> 24) & 0xFF;\n"
    "    const int vHCP = f.f2 & 0xFF, vLOS = (f.f2 >> 8) & 0xFF; const bool vBAL = ((f.f2 >> 16) & 0xFF) != 0;\n"
    "    (void)vC; (void)vD; (void)vH; (void)vS; (void)vCP; (void)vDP; (void)vHP; (void)vSP; (void)vHCP; (void)vLOS; (void)vBAL;\n";

// Compiles `prog` (which #includes the embedded device sources) for sm_100a and loads entry point `entry`.
static int jit_compile(bmc_ctx *ctx, const std::string &prog, const char *entry, JitKernel &jk, std::string &log)
{
    JitApi &api = JitApi::get();
    if (!api.ok) return fail(BMC_E_CUDA, "NVRTC specialisation unavailable: " + api.why);
    auto t0 = std::chrono::steady_clock::now();
    nvrtcProgram p = nullptr;
    if (api.CreateProgram(&p, prog.c_str(), "bmc_jit.cu", kJitSrcCount, kJitSrcTexts, kJitSrcNames) != NVRTC_SUCCESS)
        return fail(BMC_E_CUDA, "nvrtcCreateProgram failed");
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    const nvrtcResult r = api.CompileProgram(p, 3, opts);
    size_t ls = 0;
    api.GetProgramLogSize(p, &ls);
    log.assign(ls, ' ');
    if (ls) api.GetProgramLog(p, &log[0]);
    if (r != NVRTC_SUCCESS) {
        api.DestroyProgram(&p);
        return fail(BMC_E_CUDA, "NVRTC compilation failed: " + log.substr(0, 2000));
    }
    size_t cs = 0;
    api.GetCUBINSize(p, &cs);
    std::vector<char> cubin(cs);
    api.GetCUBIN(p, cubin.data());
    api.DestroyProgram(&p);
    CUresult cr = api.ModuleLoadData(&jk.module, cubin.data());
    if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "cuModuleLoadData failed: " + std::to_string((int)cr));
    cr = api.ModuleGetFunction(&jk.func, jk.module, entry);
    if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "cuModuleGetFunction failed: " + std::to_string((int)cr));
    jk.device = ctx->device;
    jk.compile_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return BMC_OK;
}

static int jit_build(bmc_ctx *ctx, bmc_constraints *c, ConsTables *tabs, bool seatfirst, bool exact)
{
    JitKernel &jk = tabs->jit[seatfirst ? (exact ? 2 : 1) : 0];      // caller holds c->cal_mu; the ctx's device is current
    if (jk.func) return BMC_OK;
    // the rule forms as one C++ function
    std::string fn = "__device__ __forceinline__ bool bmc_jit_seat_rule(int seat, const bmc::Feat &f, bool &err)\n{\n" +
                     std::string(kJitVars) + "    switch (seat) {\n";
    try {
        RuleEmitter em;
        for (int k = 0; k < 4; ++k) {
            if (c->dev.seat[k].prog == 0xFFFFFFFFu || c->rule_text[k].empty()) continue;
            SxReader rd(c->rule_text[k]);
            Sx form = rd.read();
            if (form.is_list() && form.items.size() == 2 && form.items[0].is_sym("QUOTE")) form = form.items[1];
            fn += "    case " + std::to_string(k) + ": return " + em.emit(form) + ";\n";
        }
    } catch (const RuleError &e) {
        return fail(BMC_E_PARSE, std::string("jit: ") + e.what());
    }
    fn += "    default: return true;\n    }\n}\n";
    const std::string prog = std::string(kJitPrelude) + (seatfirst ? "#define BMC_JIT_SEATFIRST 1\n" : "#define BMC_JIT_FULL 1\n") +
                             (exact ? "#define BMC_JIT_EXACT 1\n" : "") + "#include \"features.cuh\"\n" + fn + "#include \"kernels.cuh\"\n";
    int rc = jit_compile(ctx, prog, seatfirst ? "bmc_jit_sample_seatfirst" : "bmc_jit_sample_full", jk, c->jit_log);
    if (rc) return rc;
    c->jit_ms += jk.compile_ms;
    return BMC_OK;
}

// choose-bid over a rule table as one straight-line function: rows in order, first whose form holds; -2 where
// evaluation reaches an illegal form first; -1 = nil.
static int scheme_jit_build(bmc_ctx *ctx, bmc_scheme *s, SchemeTables *tabs)
{
    JitKernel &jk = tabs->jit;                        // caller holds s->mu
    if (jk.func) return BMC_OK;
    std::string fn = "__device__ __forceinline__ bool bmc_jit_seat_rule(int, const bmc::Feat &, bool &) { return true; }\n"
                     "__device__ __forceinline__ int bmc_jit_choose_bid(const bmc::Feat &f)\n{\n" + std::string(kJitVars);
    try {
        RuleEmitter em;
        for (size_t r = 0; r < s->forms.size(); ++r)
            fn += "    { bool err = false; const bool v = " + em.emit(s->forms[r]) + "; if (err) return -2; if (v) return " +
                  std::to_string(r) + "; }\n";
    } catch (const RuleError &e) {
        return fail(BMC_E_PARSE, std::string("jit: ") + e.what());
    }
    fn += "    return -1;\n}\n";
    const std::string prog = std::string(kJitPrelude) + "#define BMC_JIT_FILTER 1\n#include \"features.cuh\"\n" + fn + "#include \"filter.cuh\"\n";
    std::string log;
    int rc = jit_compile(ctx, prog, "bmc_jit_filter", jk, log);
    if (rc) return rc;
    s->jit_ms += jk.compile_ms;
    return BMC_OK;
}

int bmc_constraints_set_jit(bmc_constraints *c, int on)
{
    if (!c || on < 0 || on > 2) return fail(BMC_E_INVALID, "bmc_constraints_set_jit: bad argument");
    c->jit = on;
    return BMC_OK;
}
float bmc_constraints_jit_ms(const bmc_constraints *c)
{
    return c ? c->jit_ms : 0.f;
}

// Device-side tables of a constraint, one copy per device (a handle may be shared by contexts of several devices).
static int cons_upload(bmc_ctx *ctx, const bmc_constraints *cc, ConsDev &dev, ConsTables **tabs = nullptr)
{
    bmc_constraints *c = const_cast<bmc_constraints *>(cc);
    std::lock_guard<std::mutex> lk(c->mu);
    ConsTables &t = c->tabs[ctx->device];               // std::map: the element's address is stable
    if (tabs) *tabs = &t;
    if (!t.d_code) {
        CUDA_TRY(cudaMalloc(&t.d_code, c->code.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpyAsync(t.d_code, c->code.data(), c->code.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    dev = c->dev;
    dev.code = t.d_code;
    dev.sf_tab = reinterpret_cast<const uint4 *>(t.d_code + c->tab_off);
    return BMC_OK;
}

// ---- schemes ------------------------------------------------------------------
int bmc_scheme_create(const char *table, bmc_scheme **out)
{
    if (!out || !table) return fail(BMC_E_INVALID, "bmc_scheme_create: NULL argument");
    *out = nullptr;
    try {
        SchemeRows rows = compile_scheme(table);
        std::unique_ptr<bmc_scheme> s(new bmc_scheme());
        for (size_t i = 0; i < rows.progs.size(); ++i) {
            s->row_off.push_back((uint32_t)s->code.size());
            s->code.insert(s->code.end(), rows.progs[i].begin(), rows.progs[i].end());
        }
        s->bids = rows.bids;
        s->forms = rows.forms;
        s->code.push_back(enc(OP_END));          // spare word: the interpreter prefetches one word ahead
        if (s->row_off.size() > 120) return fail(BMC_E_CAPACITY, "rule table has more than 120 rows");
        *out = s.release();
    } catch (const RuleError &e) {
        return fail(BMC_E_PARSE, std::string("rule table: ") + e.what());
    }
    return BMC_OK;
}
int bmc_scheme_size(const bmc_scheme *s) { return s ? (int)s->row_off.size() : 0; }
void bmc_scheme_destroy(bmc_scheme *s)
{
    if (!s) return;
    for (auto &kv : s->tabs) {
        cudaSetDevice(kv.first);
        cudaFree(kv.second.d_code);
        cudaFree(kv.second.d_off);
        if (kv.second.jit.module) JitApi::get().ModuleUnload(kv.second.jit.module);
    }
    delete s;
}
int bmc_scheme_set_jit(bmc_scheme *s, int on)
{
    if (!s || on < 0 || on > 2) return fail(BMC_E_INVALID, "bmc_scheme_set_jit: bad argument");
    s->jit = on;
    return BMC_OK;
}
float bmc_scheme_jit_ms(const bmc_scheme *s) { return s ? s->jit_ms : 0.f; }
static int scheme_upload(bmc_ctx *ctx, const bmc_scheme *ss, SchemeDev &dev, SchemeTables **tabs = nullptr)
{
    dev.code = nullptr; dev.row_off = nullptr; dev.n_rows = 0;
    if (tabs) *tabs = nullptr;
    if (!ss || ss->row_off.empty()) return BMC_OK;
    bmc_scheme *s = const_cast<bmc_scheme *>(ss);
    std::lock_guard<std::mutex> lk(s->mu);
    SchemeTables &t = s->tabs[ctx->device];
    if (tabs) *tabs = &t;
    if (!t.d_code) {
        CUDA_TRY(cudaMalloc(&t.d_code, s->code.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&t.d_off, s->row_off.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpyAsync(t.d_code, s->code.data(), s->code.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(t.d_off, s->row_off.data(), s->row_off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    dev.code = t.d_code; dev.row_off = t.d_off; dev.n_rows = (uint32_t)s->row_off.size();
    return BMC_OK;
}

// ---- sampling -------------------------------------------------------------------
struct WaveCfg {                 // what is constant over the waves of one call
    ConsDev cons{};
    bool has_cons = false, seatfirst = false, exact = false, interp = false;
    const FixDev *fix = nullptr;
    const SeqDev *seq = nullptr;
    const JitKernel *jit = nullptr;
    ConsTables *tabs = nullptr;
    uint64_t seed = 0, call_first = 0, target = 0, start_acc = 0;
    uint32_t tally_mask = 0, rec_mode = 0;
    void *records = nullptr;
    uint64_t cap = 0;
    void *tallies = nullptr;
    ContractTable score{};
};

static int blocks_per_sm(const void *kernel, size_t smem)
{
    // static + dynamic shared memory may exceed the 48 KB a launch gets without asking (the reference-order kernel)
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
    static const int cap = [] { const char *e = getenv("BMC_BLOCKS_PER_SM"); return e ? atoi(e) : 0; }();      // tuning knob
    return cap > 0 && occ > cap ? cap : occ;
}

static int launch_wave(bmc_ctx *ctx, const WaveCfg &w, uint64_t first, uint64_t n)
{
    SampleArgs a;
    memset(&a, 0, sizeof a);
    a.cons = w.cons;
    if (w.fix) a.fix = *w.fix;
    else a.fix.n_free = 52;
    a.key = philox_key(w.seed);
    a.tkey = philox_key(w.seed ^ SCORE_KEY_XOR);
    a.first = first;
    a.n = n;
    a.records = w.records;
    a.cap = w.records ? w.cap : 0;
    a.tallies = (unsigned long long *)w.tallies;
    a.tally_mask = w.tally_mask;
    a.rec_mode = w.rec_mode;
    a.idx_base = (uint32_t)(first - w.call_first);
    a.flags = w.interp ? ARGS_INTERP : 0u;
    a.target = w.target;
    a.start_acc = w.start_acc;
    a.snap = ctx->d_snap;
    a.score = w.score;
    {   // stage B of the seat-first sampler: the primary seat has no capacity left
        uint32_t cap4[4] = {13, 13, 13, 13};
        cap4[w.cons.primary & 3u] = 0;
        const uint32_t p0 = cap4[0], p1 = p0 + cap4[1], p2 = p1 + cap4[2];
        a.sf_q0 = (128u - p0) | ((128u - p1) << 8) | ((128u - p2) << 16);
    }
    const bool jit = w.jit != nullptr;
    const SmemPlan plan = w.seq ? smem_plan(false, false, false, false, w.tally_mask)
                                : smem_plan(w.seatfirst, w.seatfirst || w.has_cons, w.seatfirst && sf_two_stage(w.cons, jit), w.interp, w.tally_mask);
    const size_t smem = (size_t)plan.words * 4u;
    const void *kernel = w.seq ? (const void *)sample_reforder_kernel
                       : w.seatfirst ? (w.exact ? (const void *)sample_seatfirst_kernel<true> : (const void *)sample_seatfirst_kernel<false>)
                       : w.fix ? (w.has_cons ? (const void *)sample_kernel<true, true> : (const void *)sample_kernel<false, true>)
                               : (w.has_cons ? (const void *)sample_kernel<true, false> : (const void *)sample_kernel<false, false>);
    int per_sm;
    if (jit) {
        int occ = 0;
        JitApi::get().Occupancy(&occ, w.jit->func, THREADS, smem);
        per_sm = occ > 0 ? occ : 1;
    } else per_sm = blocks_per_sm(kernel, smem);
    int blocks = ctx->sms * per_sm;
    const uint64_t per_thread = w.seatfirst ? 4 * SF_CALLS : 1;     // raw deals per thread per iteration
    const uint64_t per_iter = (uint64_t)blocks * THREADS * per_thread;
    if (n < per_iter) blocks = (int)((n + THREADS * per_thread - 1) / (THREADS * per_thread));
    if (blocks < 1) blocks = 1;
    const uint64_t chunk = (uint64_t)blocks * THREADS * per_thread;
    a.iters = (uint32_t)((n + chunk - 1) / chunk);
    if (w.seatfirst) {
        // a lane owns index QUADS (4c .. 4c + 3, one Philox call): a misaligned first index or end adds one
        const uint64_t calls = ((first + n + 3) >> 2) - (first >> 2);
        const uint64_t per = (uint64_t)blocks * THREADS * SF_CALLS;
        a.iters = (uint32_t)((calls + per - 1) / per);
    }
    if (jit) {
        void *params[] = {&a};
        const CUresult cr = JitApi::get().LaunchKernel(w.jit->func, (unsigned)blocks, 1, 1, THREADS, 1, 1, (unsigned)smem, (CUstream)ctx->stream, params, nullptr);
        if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "cuLaunchKernel (jit) failed: " + std::to_string((int)cr));
    } else if (w.seq) sample_reforder_kernel<<<blocks, THREADS, smem, ctx->stream>>>(a, *w.seq, w.tabs->root);
    else if (w.seatfirst && w.exact) sample_seatfirst_kernel<true><<<blocks, THREADS, smem, ctx->stream>>>(a);
    else if (w.seatfirst) sample_seatfirst_kernel<false><<<blocks, THREADS, smem, ctx->stream>>>(a);
    else if (w.fix) {
        if (w.has_cons) sample_kernel<true, true><<<blocks, THREADS, smem, ctx->stream>>>(a);
        else sample_kernel<false, true><<<blocks, THREADS, smem, ctx->stream>>>(a);
    } else {
        if (w.has_cons) sample_kernel<true, false><<<blocks, THREADS, smem, ctx->stream>>>(a);
        else sample_kernel<false, false><<<blocks, THREADS, smem, ctx->stream>>>(a);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return BMC_OK;
}

// ---- device distgen: class enumeration (distgen.cuh) -------------------------------------------------------
struct PatternScan {                     // host copy of pass 1
    std::vector<unsigned long long> w;   // weight per pattern
    std::vector<uint32_t> cnt;           // admitted classes per pattern
    std::vector<uint8_t> flag;           // pattern passes the honour-level tests
    unsigned long long hcp_pmf[40];
    unsigned long long len_pmf[4 * 14];
};
static int distgen_count(bmc_ctx *ctx, const ClassSpec &cs, PatternScan &ps, bool want_pmf, unsigned long long **keep_w = nullptr,
                         uint32_t **keep_cnt = nullptr)
{
    unsigned long long *d_w = nullptr, *d_pmf = nullptr;
    uint32_t *d_c = nullptr;
    uint8_t *d_f = nullptr;
    cudaError_t e = cudaMalloc(&d_w, DG_PATTERNS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&d_c, DG_PATTERNS * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_f, DG_PATTERNS);
    if (e == cudaSuccess && want_pmf) e = cudaMalloc(&d_pmf, (40 + 56) * sizeof(unsigned long long));
    if (e == cudaSuccess && want_pmf) e = cudaMemsetAsync(d_pmf, 0, (40 + 56) * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) {
        distgen_count_kernel<<<DG_PATTERNS, DG_THREADS, 0, ctx->stream>>>(cs, d_w, d_c, d_f, d_pmf, d_pmf ? d_pmf + 40 : nullptr);
        ctx->launches++;
        e = cudaGetLastError();
    }
    ps.w.resize(DG_PATTERNS); ps.cnt.resize(DG_PATTERNS); ps.flag.resize(DG_PATTERNS);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ps.w.data(), d_w, DG_PATTERNS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ps.cnt.data(), d_c, DG_PATTERNS * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ps.flag.data(), d_f, DG_PATTERNS, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && want_pmf) {
        e = cudaMemcpyAsync(ps.hcp_pmf, d_pmf, 40 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ps.len_pmf, d_pmf + 40, 56 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_f); cudaFree(d_pmf);
    if (keep_w && e == cudaSuccess) *keep_w = d_w; else cudaFree(d_w);
    if (keep_cnt && e == cudaSuccess) *keep_cnt = d_c; else cudaFree(d_c);
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("distgen: ") + cudaGetErrorString(e));
    return BMC_OK;
}

constexpr uint64_t EXACT_MAX_CLASSES = 1ull << 23;     // all 4 582 640 classes of a 13-card hand fit (55 MB of tables)

// Enumerates the classes `cs` admits and, when there are between 1 and EXACT_MAX_CLASSES of them, leaves their table on
// the device in t->x_*.  scale = 0: the rank of u is u / floor(2^64 / W) (a generator of its own: mode 3's root);
// otherwise u / scale (mode 4: scale = floor(2^64 / C(52,13)), the admitted hands come first).
static int class_table_build(bmc_ctx *ctx, const ClassSpec &cs, unsigned long long scale, ConsTables *t, unsigned long long *W_out,
                             unsigned long long *K_out, bool *built)
{
    *built = false;
    PatternScan ps;
    unsigned long long *d_w = nullptr;
    uint32_t *d_c = nullptr;
    { int rc = distgen_count(ctx, cs, ps, false, &d_w, &d_c); if (rc) return rc; }
    // exclusive scans over the 65 536 patterns (host: a control-plane step, once per constraint and device)
    std::vector<unsigned long long> pcum(DG_PATTERNS);
    std::vector<uint32_t> poff(DG_PATTERNS);
    unsigned long long W = 0, K = 0;
    for (int p = 0; p < DG_PATTERNS; ++p) { pcum[p] = W; poff[p] = (uint32_t)K; W += ps.w[p]; K += ps.cnt[p]; }
    *W_out = W; *K_out = K;
    if (K == 0 || K > EXACT_MAX_CLASSES) { cudaFree(d_w); cudaFree(d_c); return BMC_OK; }
    const unsigned long long S = scale ? scale : ~0ull / W;                 // floor((2^64 - 1) / W): fits 64 bits even for W = 1
    unsigned long long *d_pcum = nullptr;
    uint32_t *d_poff = nullptr;
    cudaError_t e = cudaMalloc(&d_pcum, DG_PATTERNS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&d_poff, DG_PATTERNS * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&t->x_cum, (K + 1) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&t->x_keys, K * sizeof(uint32_t));
    // buckets: about two per class, a power of two in [2^8, 2^23]
    uint32_t nb = 256;
    while (nb < 2 * K && nb < (1u << 23)) nb <<= 1;
    const unsigned long long theta = W * S;                  // <= 2^64 - 1
    unsigned long long bm = ((unsigned long long)nb << 32) / ((theta >> 32) + 1ull);
    const uint32_t bmul = bm > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)bm;
    if (e == cudaSuccess) e = cudaMalloc(&t->x_bkt, ((size_t)nb + 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_pcum, pcum.data(), DG_PATTERNS * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_poff, poff.data(), DG_PATTERNS * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->x_cum + K, &W, sizeof W, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        distgen_emit_kernel<<<DG_PATTERNS, DG_THREADS, 0, ctx->stream>>>(cs, d_pcum, d_poff, d_c, t->x_keys, t->x_cum);
        distgen_bucket_kernel<<<(nb + 256) / 256, 256, 0, ctx->stream>>>(t->x_cum, (uint32_t)K, bmul, nb, S, t->x_bkt);
        ctx->launches += 2;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_w); cudaFree(d_c); cudaFree(d_pcum); cudaFree(d_poff);
    if (e != cudaSuccess) {
        cudaFree(t->x_cum); cudaFree(t->x_keys); cudaFree(t->x_bkt);
        t->x_cum = nullptr; t->x_keys = nullptr; t->x_bkt = nullptr;
        return fail(BMC_E_CUDA, std::string("class table: ") + cudaGetErrorString(e));
    }
    t->x_bmul = bmul; t->x_nbkt = nb; t->x_built = true;
    *built = true;
    return BMC_OK;
}

// The exact class table of the primary seat on this device (mode 4).  Always leaves the exact acceptance of the seat in
// the constraint (x_weight / C(52,13)); *built = false when nothing is admitted.
static int exact_build(bmc_ctx *ctx, bmc_constraints *c, ConsTables *t, bool *built)
{
    *built = false;
    if (t->x_built) { *built = true; return BMC_OK; }
    const uint32_t P = c->dev.primary;
    ClassSpec cs{};
    cs.seat = c->dev.seat[P];
    cs.prog = cs.seat.prog != 0xFFFFFFFFu ? t->d_code + cs.seat.prog : nullptr;
    for (int s = 0; s < 4; ++s) cs.n_low[s] = 9;
    cs.hon_avail = 0xFFFFu;
    cs.hand_size = 13;
    unsigned long long W = 0, K = 0;
    int rc = class_table_build(ctx, cs, SF_S, t, &W, &K, built);
    if (rc) return rc;
    c->x_weight = W; c->x_classes = K; c->x_known = true;
    c->stage_a_pass = (double)W / (double)SF_HANDS;
    return BMC_OK;
}

static void ranges_to_seatc(const SeatRanges &r, SeatC &c);

// Mode 3: the root generator (first ranged seat over the deck the fixed hands leave) as a class table on this device.
static int root_build(bmc_ctx *ctx, bmc_constraints *c, ConsTables *t)
{
    if (t->x_built || t->root.limit) return BMC_OK;          // built, or known to be empty (limit = 1 marks it)
    const SeqDev &q = c->seq;
    const SeqRanges r = seq_infer(q.st[0], q.exhaust, c->fix.free0, c->fix.free1);
    SeqRoot root{};
    root.low0 = c->fix.free0 & 0x01FF01FFu;
    root.low1 = c->fix.free1 & 0x01FF01FFu;
    root.full = (root.low0 == 0x01FF01FFu && root.low1 == 0x01FF01FFu) ? 1u : 0u;
    ClassSpec cs{};
    cs.hand_size = 13;
    for (int s = 0; s < 4; ++s) {
        const uint32_t lane = ((s < 2 ? c->fix.free0 : c->fix.free1) >> (16 * (s & 1))) & 0x1FFFu;
        cs.n_low[s] = (uint32_t)__builtin_popcount(lane & 0x1FFu);
        for (int h = 0; h < 4; ++h)
            if ((lane >> (9 + h)) & 1u) cs.hon_avail |= 1u << (15 - (4 * s + h));
    }
    SeatRanges sr;
    bool empty = r.bad;
    sr.lo[F_HCP] = r.hcp_lo < 0 ? 0 : r.hcp_lo; sr.hi[F_HCP] = r.hcp_hi > 40 ? 40 : r.hcp_hi;
    for (int s = 0; s < 4; ++s) {
        sr.lo[F_C + s] = r.len_lo[s] < 0 ? 0 : r.len_lo[s]; sr.hi[F_C + s] = r.len_hi[s] > 13 ? 13 : r.len_hi[s];
        sr.lo[F_CP + s] = r.sh_lo[s] < 0 ? 0 : r.sh_lo[s]; sr.hi[F_CP + s] = r.sh_hi[s] > 10 ? 10 : r.sh_hi[s];
    }
    empty = empty || sr.empty();
    unsigned long long W = 0, K = 0;
    bool built = false;
    if (!empty) {
        SeatC sc{};
        ranges_to_seatc(sr, sc);
        sc.prog = 0xFFFFFFFFu; sc.active = 1;
        cs.seat = sc;
        cs.prog = nullptr;
        int rc = class_table_build(ctx, cs, 0ull, t, &W, &K, &built);
        if (rc) return rc;
    }
    if (built) {
        root.W = W;
        root.S = ~0ull / W;                                  // every hand has exactly S preimages among the u < W S
        root.S_recip = (unsigned long long)((((unsigned __int128)1) << 64) / root.S);
        root.limit = W * root.S;
        root.x.cum = t->x_cum; root.x.keys = t->x_keys; root.x.bkt = t->x_bkt; root.x.bmul = t->x_bmul; root.x.n_bkt = t->x_nbkt;
    } else {
        root.W = 0;                                          // the root generator has total weight 0: every build signals an error
        root.limit = 1;
    }
    t->root = root;
    return BMC_OK;
}

// Resolves how `cons` is sampled on this context's device: uploads its tables, calibrates the dealing mode once per
// constraint (reproducibly: fixed calibration seed), builds the NVRTC kernel when asked.  Fills the constant part of `w`.
static int resolve_cfg(bmc_ctx *ctx, const bmc_constraints *cons, WaveCfg &w)
{
    w.has_cons = cons && cons->has_cons;
    if (cons) { int rc = cons_upload(ctx, cons, w.cons, &w.tabs); if (rc) return rc; }
    if (!w.has_cons) w.cons = ConsDev{};
    w.fix = (cons && cons->has_fixed) ? &cons->fix : nullptr;
    if (cons) w.score = cons->score;
    if (cons && cons->resolved == 3) {                 // reference order: `build`'s own sequential semantics
        bmc_constraints *cc = const_cast<bmc_constraints *>(cons);
        std::lock_guard<std::mutex> cal_lock(cc->cal_mu);
        int rc = root_build(ctx, cc, w.tabs);
        if (rc) return rc;
        w.seq = &cons->seq;
        w.fix = &cons->fix;
    } else if (w.has_cons && !w.fix) {
        bmc_constraints *cc = const_cast<bmc_constraints *>(cons);
        std::lock_guard<std::mutex> cal_lock(cc->cal_mu);      // one calibration per constraint, even across contexts
        // pass rate of `cal` over 2^16 raw deals of the full-deal kernel
        auto pass_rate = [&](const ConsDev &cal, double &rate) -> int {
            unsigned long long *d_cal = nullptr, h4[4] = {0, 0, 0, 0};
            CUDA_TRY(cudaMalloc(&d_cal, sizeof(bmc_tallies)));
            cudaMemsetAsync(d_cal, 0, sizeof(bmc_tallies), ctx->stream);
            WaveCfg wc;
            wc.cons = cal; wc.has_cons = true; wc.seed = 0x63616C6962726174ull; wc.tallies = d_cal;
            int rc = launch_wave(ctx, wc, 0, 1u << 16);
            cudaError_t e = cudaMemcpyAsync(h4, d_cal, sizeof h4, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_cal);
            if (rc) return rc;
            if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("calibration: ") + cudaGetErrorString(e));
            rate = (double)h4[2] / (double)(h4[0] ? h4[0] : 1);
            return BMC_OK;
        };
        if (cc->resolved == 0) {
            // The device enumeration (distgen.cuh) gives the primary seat's EXACT acceptance -- no sampling
            // calibration -- and, when it is small enough, the class table of mode 4.
            bool built = false;
            if (cc->mode == 0 || cc->mode == 4) { int rc = exact_build(ctx, cc, w.tabs, &built); if (rc) return rc; }
            if ((cc->mode == 0 || cc->mode == 4) && cc->x_weight == 0)
                return fail(BMC_E_BAD_RANGE, "seat " + std::to_string(w.cons.primary) + ": no 13-card hand satisfies its constraint (total weight 0)");
            if (cc->mode == 4 && !built)
                return fail(BMC_E_CAPACITY, "mode 4: the primary seat admits " + std::to_string(cc->x_classes) + " classes of hands, more than the table holds");
            if (cc->mode != 0) cc->resolved = cc->mode;
            else cc->resolved = cc->stage_a_pass >= 0.30 ? 1 : built ? 4 : 2;      // seat-first pays off when most deals die on the primary seat
            cc->dev.b1_mask = 0xFu;
            if ((cc->resolved == 2 || cc->resolved == 4) && w.cons.n_active >= 3) {
                // the most selective other seat: tested together with the primary seat by the compiled-rule kernel
                double best = 2.0;
                int second = -1;
                for (int k = 0; k < 4; ++k) {
                    if ((uint32_t)k == w.cons.primary || !w.cons.seat[k].active) continue;
                    ConsDev cal = w.cons;
                    for (int j = 0; j < 4; ++j) if (j != k) cal.seat[j].active = 0;
                    cal.primary = (uint32_t)k;
                    cal.n_prog = cal.seat[k].prog != 0xFFFFFFFFu ? 1u : 0u;
                    double r = 1.0;
                    int rc = pass_rate(cal, r);
                    if (rc) return rc;
                    if (r < best) { best = r; second = k; }
                }
                if (second >= 0) cc->dev.b1_mask = (1u << w.cons.primary) | (1u << second);
            }
        }
        w.cons.b1_mask = cc->dev.b1_mask;
        w.seatfirst = cc->resolved == 2 || cc->resolved == 4;
        w.exact = cc->resolved == 4;
        if (w.exact) {
            bool built = false;
            int rc = exact_build(ctx, cc, w.tabs, &built);             // per device; a no-op once built
            if (rc) return rc;
            if (!built) return fail(BMC_E_CAPACITY, "mode 4: class table unavailable");
            w.cons.x.cum = w.tabs->x_cum; w.cons.x.keys = w.tabs->x_keys; w.cons.x.bkt = w.tabs->x_bkt;
            w.cons.x.bmul = w.tabs->x_bmul; w.cons.x.n_bkt = w.tabs->x_nbkt;
            w.cons.sf_theta = cc->x_weight * SF_S;
            // the primary seat is admitted by construction: its program never runs in the sampling kernel
            if (w.cons.seat[w.cons.primary].prog != 0xFFFFFFFFu && w.cons.n_prog) w.cons.n_prog--;
        }
    }
    w.interp = w.has_cons && w.cons.n_prog != 0u;
    if (cons && cons->has_rules && w.has_cons && !w.fix && !w.seq) {
        bmc_constraints *cc = const_cast<bmc_constraints *>(cons);
        std::lock_guard<std::mutex> jl(cc->cal_mu);
        // jit 1: always; jit 2 (default): once the constraint has spent more than a second of kernel time interpreting
        // its rule programs -- about what the NVRTC build costs.  Results are bit-identical either way.
        const bool want = cc->jit == 1 || (cc->jit == 2 && cc->interp_ms > 1000.0 && w.cons.n_prog != 0u && JitApi::get().ok);
        if (want) {
            int rc = jit_build(ctx, cc, w.tabs, w.seatfirst, w.exact);
            if (rc) {
                if (cc->jit == 1) return rc;
                cc->jit = 0;                           // auto: the build failed once -- keep interpreting
            } else {
                w.jit = &w.tabs->jit[w.seatfirst ? (w.exact ? 2 : 1) : 0];
                w.interp = false;
            }
        }
    }
    return BMC_OK;
}

static int publish(bmc_ctx *ctx, const void *tallies_dev, unsigned long long seq, const unsigned long long *extra = nullptr)
{
    publish_counters_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)tallies_dev, ctx->h_head_dev, ctx->d_snap, seq, extra);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return BMC_OK;
}

// packed index records (pack.cuh): the waves' 32-bit offsets (rec_mode 1, `records_dev`) are turned into gap bytes wave
// by wave, and the BYTES are what is streamed to the host
struct PackCfg { uint8_t *bytes_dev; uint64_t cap_bytes; uint8_t *host_bytes; uint64_t *n_bytes; };

// Common driver.  An accept-target job enqueues its waves back to back in batches sized from the acceptance seen so
// far: a wave that starts after the target was met cancels itself on the device (wave_cancelled), so the accepted
// set is still "whole waves from first_index until the target is met" while the host synchronises once per batch
// instead of once per wave.  host_out != NULL: records are streamed to the caller's buffer on a second stream while
// later waves compute; the host follows the waves through the counters each one publishes in mapped memory.
static int sample_core(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                       uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, void *records_dev, uint64_t cap,
                       void *tallies_dev, bmc_stats *stats, void *host_out, uint32_t rec_mode, const PackCfg *pack = nullptr)
{
    if (n_accept_target == 0 && n_raw == 0) return fail(BMC_E_INVALID, "nothing to do: n_raw == 0 and no accept target");
    if ((tally_mask & BMC_TALLY_SCORE) && (!cons || cons->score.n == 0))
        return fail(BMC_E_INVALID, "BMC_TALLY_SCORE needs contracts: bmc_constraints_set_scoring");
    auto t0 = std::chrono::steady_clock::now();
    WaveCfg w;
    { int rc = resolve_cfg(ctx, cons, w); if (rc) return rc; }
    w.seed = seed; w.call_first = first_index; w.tally_mask = tally_mask; w.rec_mode = rec_mode;
    w.records = records_dev; w.cap = cap; w.tallies = tallies_dev;
    const uint64_t elem = rec_mode ? sizeof(uint32_t) : sizeof(bmc_record);
    if (rec_mode) {
        // index records are 32-bit offsets from first_index: a call spans at most 2^32 indices
        if (n_raw > (1ull << 32)) return fail(BMC_E_INVALID, "index records: at most 2^32 indices per call");
        if (n_raw == 0) n_raw = 1ull << 32;
    }
    const bool stream_copy = pack ? true : (host_out && records_dev && cap);
    const bool stepwise = n_accept_target != 0 || stream_copy;
    if (wave == 0) wave = 1ull << 26;
    // 32-bit offsets from the launch's first index travel through the kernels' queues: 2^32 indices per launch
    // (2^31 for the seat-first kernel, whose offsets are computed in quads)
    uint64_t max_launch = stepwise ? wave : (1ull << 32);
    if (max_launch > (1ull << 32)) max_launch = 1ull << 32;
    if (w.seatfirst && max_launch > (1ull << 31)) max_launch = 1ull << 31;
    const uint64_t launches0 = ctx->launches;
    volatile unsigned long long *head = ctx->h_head;
    for (int i = 0; i < 5; ++i) head[i] = 0;
    unsigned long long seq = 0;
    const unsigned long long *cursor = pack ? &ctx->d_pack->bytes : nullptr;
    auto sync_head = [&]() -> int {                      // publish, wait for the stream, the counters are in `head`
        int rc = publish(ctx, tallies_dev, ++seq, cursor);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return BMC_OK;
    };
    { int rc = sync_head(); if (rc) return rc; }         // counters already in the caller's tally buffer
    const unsigned long long start_raw = head[0], start_acc = head[2];
    w.target = n_accept_target;
    w.start_acc = start_acc;
    uint64_t copied = start_acc < cap ? start_acc : cap;
    if (pack) {
        copied = 0;
        const PackState init = {start_acc < cap ? start_acc : cap, 0ull, 0ull, 0ull};
        CUDA_TRY(cudaMemcpyAsync(ctx->d_pack, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    }
    // after a wave: its new offsets -> bitmap -> gap bytes appended to the byte stream
    auto pack_wave = [&](uint64_t wave_done, uint64_t n) -> int {
        const uint64_t words = (n + 31) / 32;
        CUDA_TRY(cudaMemsetAsync(ctx->d_bitmap, 0, words * 4, ctx->stream));
        pack_scatter_kernel<<<ctx->sms * 8, 256, 0, ctx->stream>>>((const uint32_t *)records_dev, (const unsigned long long *)tallies_dev,
                                                                   ctx->d_pack, cap, (uint32_t)wave_done, (uint32_t)n, ctx->d_bitmap);
        pack_tiles_kernel<<<(unsigned)((n + PACK_TILE_BITS - 1) / PACK_TILE_BITS), PACK_THREADS, 0, ctx->stream>>>(
            ctx->d_bitmap, (uint32_t)wave_done, (uint32_t)n, pack->bytes_dev, pack->cap_bytes, ctx->d_pack);
        pack_advance_kernel<<<1, 1, 0, ctx->stream>>>((const unsigned long long *)tallies_dev, ctx->d_pack, cap);
        ctx->launches += 3;
        CUDA_TRY(cudaGetLastError());
        return BMC_OK;
    };
    auto copy_upto = [&](unsigned long long accepted) -> int {
        if (pack) {                                      // bytes, not records: the cursor travels in head[5]
            unsigned long long upto = head[5];
            if (upto > pack->cap_bytes) upto = pack->cap_bytes;
            if (upto > copied) {
                CUDA_TRY(cudaMemcpyAsync(pack->host_bytes + copied, pack->bytes_dev + copied, upto - copied, cudaMemcpyDeviceToHost, ctx->copy_stream));
                copied = upto;
            }
            return BMC_OK;
        }
        const uint64_t upto = accepted < cap ? accepted : cap;
        if (stream_copy && upto > copied) {
            CUDA_TRY(cudaMemcpyAsync((char *)host_out + copied * elem, (const char *)records_dev + copied * elem,
                                     (upto - copied) * elem, cudaMemcpyDeviceToHost, ctx->copy_stream));
            copied = upto;
        }
        return BMC_OK;
    };
    float kernel_ms = 0.f;
    uint64_t done = 0;                                    // indices enqueued (cancelled waves included)
    bmc_constraints *cc = const_cast<bmc_constraints *>(cons);
    if (!stepwise) {
        CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        while (done < n_raw) {
            const uint64_t n = n_raw - done < max_launch ? n_raw - done : max_launch;
            int rc = launch_wave(ctx, w, first_index + done, n);
            if (rc) return rc;
            done += n;
        }
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        { int rc = sync_head(); if (rc) return rc; }
        cudaEventElapsedTime(&kernel_ms, ctx->ev0, ctx->ev1);
    } else {
        uint64_t batch = 1;
        for (;;) {
            if (n_raw && done >= n_raw) break;
            CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
            const unsigned long long seq0 = seq;
            for (uint64_t k = 0; k < batch && !(n_raw && done >= n_raw); ++k) {
                uint64_t n = max_launch;
                if (n_raw && n_raw - done < n) n = n_raw - done;
                int rc = launch_wave(ctx, w, first_index + done, n);
                if (rc) return rc;
                if (pack) { rc = pack_wave(done, n); if (rc) return rc; }
                rc = publish(ctx, tallies_dev, ++seq, cursor);
                if (rc) return rc;
                done += n;
            }
            CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
            if (stream_copy) {
                // follow the waves of the batch as they finish and copy what each one added
                unsigned long long seen = seq0;
                unsigned spins = 0;
                while (seen < seq) {
                    const unsigned long long now = head[4];
                    if (now > seen) {
                        seen = now;
                        std::atomic_thread_fence(std::memory_order_acquire);
                        int rc = copy_upto(head[2]);
                        if (rc) return rc;
                    } else if ((++spins & 0x3FFu) == 0u) {
                        const cudaError_t q = cudaStreamQuery(ctx->stream);
                        if (q == cudaSuccess) break;                         // drained: the final read below sees everything
                        if (q != cudaErrorNotReady) return fail(BMC_E_CUDA, std::string("sampling kernel: ") + cudaGetErrorString(q));
                        std::this_thread::yield();
                    }
                }
            }
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            { int rc = copy_upto(head[2]); if (rc) return rc; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
            kernel_ms += ms;
            if (cc && w.interp) { std::lock_guard<std::mutex> jl(cc->cal_mu); cc->interp_ms += ms; }
            const unsigned long long acc = head[2] - start_acc, raw = head[0] - start_raw;
            if (n_accept_target && acc >= n_accept_target) break;
            // watchdog for open-ended jobs: a constraint nothing satisfies must not spin forever
            if (n_raw == 0 && ((done >= (1ull << 36) && acc == 0) || done >= (1ull << 44))) break;
            if (w.seq && acc == 0 && head[1] >= done) break;                 // reference order: every build signalled an error
            if (!n_accept_target) batch = 8;
            else if (acc == 0 || raw == 0) batch = batch < 32 ? batch * 2 : 64;
            else {
                const double per_wave = (double)acc / (double)raw * (double)max_launch;
                const double need = (double)(n_accept_target - acc) / per_wave;
                batch = need < 1.0 ? 1 : need > 64.0 ? 64 : (uint64_t)(need + 0.999);
            }
            // a long job on interpreted rules: switch to the compiled kernel between batches (bit-identical results)
            if (cc && w.interp && cc->jit == 2 && cc->interp_ms > 1000.0) {
                int rc = resolve_cfg(ctx, cons, w);
                if (rc) return rc;
            }
        }
    }
    if (!stepwise && cc && w.interp) { std::lock_guard<std::mutex> jl(cc->cal_mu); cc->interp_ms += kernel_ms; }
    if (stream_copy) CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    if (pack) {
        PackState fin;
        CUDA_TRY(cudaMemcpyAsync(&fin, ctx->d_pack, sizeof fin, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (pack->n_bytes) *pack->n_bytes = fin.bytes;
        if (fin.overflow || fin.bytes > pack->cap_bytes || head[2] > cap)
            return fail(BMC_E_CAPACITY, "bmc_sample_packed: the byte buffer (or the internal offset scratch) is too small; *n_bytes = bytes needed");
    }
    if (stats) {
        stats->raw = head[0] - start_raw; stats->voided = head[1]; stats->accepted = head[2];
        stats->stored = records_dev ? (head[2] < cap ? head[2] : cap) : 0;
        stats->waves = (uint32_t)((stats->raw + max_launch - 1) / max_launch);        // waves that ran (cancelled ones draw nothing)
        stats->launches = (uint32_t)(ctx->launches - launches0);
        stats->kernel_ms = kernel_ms;
        stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return BMC_OK;
}

static int ensure_scratch(bmc_ctx *ctx, uint64_t bytes)
{
    // device scratch for the records is kept by the context and only ever grows
    if (bytes > ctx->rec_bytes) {
        if (ctx->d_rec) cudaFree(ctx->d_rec);
        ctx->d_rec = nullptr;
        ctx->rec_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_rec, bytes));
        ctx->rec_bytes = bytes;
    }
    return BMC_OK;
}

static int sample_host(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                       uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, void *records, uint64_t cap,
                       bmc_tallies *tallies, bmc_stats *stats, uint32_t rec_mode)
{
    if (!ctx) return fail(BMC_E_INVALID, "ctx is NULL");
    auto t0 = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!records) cap = 0;
    { int rc = ensure_scratch(ctx, cap * (rec_mode ? sizeof(uint32_t) : sizeof(bmc_record))); if (rc) return rc; }
    CUDA_TRY(cudaMemsetAsync(ctx->d_tallies, 0, sizeof(bmc_tallies), ctx->stream));
    bmc_stats st{};
    int rc = sample_core(ctx, cons, seed, first_index, n_raw, n_accept_target, wave, tally_mask, cap ? ctx->d_rec : nullptr, cap,
                         ctx->d_tallies, &st, cap ? records : nullptr, rec_mode);
    if (rc != BMC_OK) return rc;
    bmc_tallies host_t;
    CUDA_TRY(cudaMemcpyAsync(&host_t, ctx->d_tallies, sizeof host_t, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    host_t.stored = st.stored;
    if (tallies) *tallies = host_t;
    if (stats) {
        *stats = st;
        stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return BMC_OK;
}

int bmc_sample_dev(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                   uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, void *records_dev, uint64_t cap,
                   void *tallies_dev, bmc_stats *stats)
{
    if (!ctx || !tallies_dev) return fail(BMC_E_INVALID, "bmc_sample_dev: ctx / tallies_dev is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return sample_core(ctx, cons, seed, first_index, n_raw, n_accept_target, wave, tally_mask, records_dev, cap, tallies_dev,
                       stats, nullptr, 0u);
}

int bmc_sample(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
               uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, bmc_record *records, uint64_t cap,
               bmc_tallies *tallies, bmc_stats *stats)
{
    return sample_host(ctx, cons, seed, first_index, n_raw, n_accept_target, wave, tally_mask, records, cap, tallies, stats, 0u);
}

// ---- index records: 4 bytes per accepted deal instead of 16, expanded on demand -------------------------------
int bmc_sample_indices(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                       uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, uint32_t *offsets, uint64_t cap,
                       bmc_tallies *tallies, bmc_stats *stats)
{
    return sample_host(ctx, cons, seed, first_index, n_raw, n_accept_target, wave, tally_mask, offsets, cap, tallies, stats, 1u);
}

int bmc_sample_indices_dev(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                           uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, void *offsets_dev, uint64_t cap,
                           void *tallies_dev, bmc_stats *stats)
{
    if (!ctx || !tallies_dev) return fail(BMC_E_INVALID, "bmc_sample_indices_dev: ctx / tallies_dev is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return sample_core(ctx, cons, seed, first_index, n_raw, n_accept_target, wave, tally_mask, offsets_dev, cap, tallies_dev,
                       stats, nullptr, 1u);
}

// ---- packed index records: about one byte per accepted deal (pack.cuh) ------------------------------------------
int bmc_sample_packed(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                      uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, uint8_t *bytes, uint64_t cap_bytes,
                      uint64_t *n_bytes, bmc_tallies *tallies, bmc_stats *stats)
{
    if (!ctx || !bytes || cap_bytes == 0) return fail(BMC_E_INVALID, "bmc_sample_packed: ctx / bytes is NULL");
    auto t0 = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (wave == 0) wave = 1ull << 26;
    if (wave > (1ull << 31)) wave = 1ull << 31;
    // scratch: the call's offsets (a byte of the stream stands for at least one deal: cap_bytes offsets suffice), one
    // wave's bitmap, the byte stream, the cursors
    const uint64_t cap_off = cap_bytes;
    { int rc = ensure_scratch(ctx, cap_off * sizeof(uint32_t)); if (rc) return rc; }
    if (wave / 8 + 4 > ctx->bitmap_bytes) {
        if (ctx->d_bitmap) cudaFree(ctx->d_bitmap);
        ctx->d_bitmap = nullptr; ctx->bitmap_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_bitmap, wave / 8 + 4));
        ctx->bitmap_bytes = wave / 8 + 4;
    }
    if (cap_bytes > ctx->packed_bytes) {
        if (ctx->d_packed) cudaFree(ctx->d_packed);
        ctx->d_packed = nullptr; ctx->packed_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_packed, cap_bytes));
        ctx->packed_bytes = cap_bytes;
    }
    if (!ctx->d_pack) CUDA_TRY(cudaMalloc(&ctx->d_pack, sizeof(PackState)));
    CUDA_TRY(cudaMemsetAsync(ctx->d_tallies, 0, sizeof(bmc_tallies), ctx->stream));
    bmc_stats st{};
    uint64_t nb = 0;
    const PackCfg pack = {ctx->d_packed, cap_bytes, bytes, &nb};
    const int rc = sample_core(ctx, cons, seed, first_index, n_raw, n_accept_target, wave, tally_mask, ctx->d_rec, cap_off,
                               ctx->d_tallies, &st, nullptr, 1u, &pack);
    if (n_bytes) *n_bytes = nb;
    if (rc != BMC_OK) return rc;
    bmc_tallies host_t;
    CUDA_TRY(cudaMemcpyAsync(&host_t, ctx->d_tallies, sizeof host_t, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    host_t.stored = st.accepted;
    if (tallies) *tallies = host_t;
    if (stats) {
        *stats = st;
        stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return BMC_OK;
}

// the byte stream -> 32-bit offsets from the call's first index (host side: a format decoder, no sampling)
int bmc_unpack_indices(const uint8_t *bytes, uint64_t n_bytes, uint32_t *offsets, uint64_t cap, uint64_t *n)
{
    if (!bytes || !n) return fail(BMC_E_INVALID, "bmc_unpack_indices: NULL argument");
    uint64_t pos = 0, count = 0;
    while (pos + 8 <= n_bytes) {
        uint32_t base = 0, len = 0;
        for (int i = 0; i < 4; ++i) { base |= (uint32_t)bytes[pos + i] << (8 * i); len |= (uint32_t)bytes[pos + 4 + i] << (8 * i); }
        pos += 8;
        if (pos + len > n_bytes) return fail(BMC_E_INVALID, "bmc_unpack_indices: truncated chunk");
        uint64_t cur = (uint64_t)base;                   // one past the "index before the tile"
        for (uint32_t k = 0; k < len; ++k) {
            const uint8_t b = bytes[pos + k];
            if (b == 255) { cur += 255; continue; }
            cur += b;
            if (offsets && count < cap) offsets[count] = (uint32_t)cur;
            ++count;
            cur += 1;
        }
        pos += len;
    }
    *n = count;
    if (pos != n_bytes) return fail(BMC_E_INVALID, "bmc_unpack_indices: trailing bytes");
    if (offsets && count > cap) return fail(BMC_E_CAPACITY, "bmc_unpack_indices: more indices than cap");
    return BMC_OK;
}

static int expand_core(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, const uint32_t *d_off,
                       uint64_t n, uint4 *d_out)
{
    WaveCfg w;
    { int rc = resolve_cfg(ctx, cons, w); if (rc) return rc; }
    SampleArgs a;
    memset(&a, 0, sizeof a);
    a.cons = w.cons;
    if (w.fix) a.fix = *w.fix;
    else a.fix.n_free = 52;
    a.key = philox_key(seed);
    a.first = first_index;
    {
        uint32_t cap4[4] = {13, 13, 13, 13};
        cap4[w.cons.primary & 3u] = 0;
        const uint32_t p0 = cap4[0], p1 = p0 + cap4[1], p2 = p1 + cap4[2];
        a.sf_q0 = (128u - p0) | ((128u - p1) << 8) | ((128u - p2) << 16);
    }
    const int mode = w.seq ? 3 : w.seatfirst ? (w.exact ? 4 : 2) : 1;
    static const SeqDev no_seq{};
    static const SeqRoot no_root{};
    uint64_t want = (n + THREADS - 1) / THREADS;
    const unsigned blocks = (unsigned)(want < (uint64_t)ctx->sms * 8 ? want : (uint64_t)ctx->sms * 8);
    expand_kernel<<<blocks, THREADS, 0, ctx->stream>>>(a, w.seq ? *w.seq : no_seq, w.seq ? w.tabs->root : no_root, mode, d_off, n, d_out);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return BMC_OK;
}

// records[i] = the deal behind index first_index + offsets[i], for the same (cons, seed) the offsets were sampled
// with (bmc_sample_indices).  Host buffers.
int bmc_expand(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, const uint32_t *offsets,
               uint64_t n, bmc_record *records)
{
    if (!ctx) return fail(BMC_E_INVALID, "ctx is NULL");
    if (n == 0) return BMC_OK;
    if (!offsets || !records) return fail(BMC_E_INVALID, "bmc_expand: NULL buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t chunk = 1ull << 24;                   // bounded device footprint: 16 Mi deals per pass
    const uint64_t m = n < chunk ? n : chunk;
    { int rc = ensure_scratch(ctx, m * (sizeof(uint32_t) + sizeof(bmc_record))); if (rc) return rc; }
    uint4 *d_out = (uint4 *)ctx->d_rec;
    uint32_t *d_off = (uint32_t *)((char *)ctx->d_rec + m * sizeof(bmc_record));
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t k = n - off < chunk ? n - off : chunk;
        CUDA_TRY(cudaMemcpyAsync(d_off, offsets + off, k * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        int rc = expand_core(ctx, cons, seed, first_index, d_off, k, d_out);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(records + off, d_out, k * sizeof(bmc_record), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return BMC_OK;
}

int bmc_expand_dev(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, const void *offsets_dev,
                   uint64_t n, void *records_dev)
{
    if (!ctx) return fail(BMC_E_INVALID, "ctx is NULL");
    if (n == 0) return BMC_OK;
    if (!offsets_dev || !records_dev) return fail(BMC_E_INVALID, "bmc_expand_dev: NULL buffer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return expand_core(ctx, cons, seed, first_index, (const uint32_t *)offsets_dev, n, (uint4 *)records_dev);
}

// ---- one call, several devices ---------------------------------------------------------------------------------
// The index range [first_index, first_index + n_raw) is split into contiguous shares, one host thread and one context
// per device; the only exchange is the sum of the tally buffers (device 0 reads its peers' over NVLink when peer
// access is available, else the host adds them).  Because a deal is a pure function of (seed, index), tallies and the
// accepted set are identical for every number of devices.  With an accept target every device runs whole waves of its
// own sub-range [first_index + g 2^40, ...) until it holds its share of the target.
struct MultiPool {
    std::mutex mu;
    std::map<int, bmc_ctx *> ctx;
    ~MultiPool() { for (auto &kv : ctx) bmc_shutdown(kv.second); }
};
static MultiPool g_pool;

int bmc_sample_multi(const int *devices, int n_devices, const bmc_constraints *cons, uint64_t seed, uint64_t first_index,
                     uint64_t n_raw, uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, bmc_record *records,
                     uint64_t cap, bmc_tallies *tallies, bmc_stats *stats)
{
    if (!devices || n_devices < 1 || n_devices > 16) return fail(BMC_E_INVALID, "bmc_sample_multi: 1..16 devices");
    if (n_accept_target == 0 && n_raw == 0) return fail(BMC_E_INVALID, "nothing to do: n_raw == 0 and no accept target");
    auto t0 = std::chrono::steady_clock::now();
    const int G = n_devices;
    std::vector<bmc_ctx *> ctxs(G, nullptr);
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        for (int g = 0; g < G; ++g) {
            for (int h = 0; h < g; ++h)
                if (devices[h] == devices[g]) return fail(BMC_E_INVALID, "bmc_sample_multi: a device is listed twice");
            bmc_ctx *&c = g_pool.ctx[devices[g]];
            if (!c) { int rc = bmc_init(devices[g], &c); if (rc) { g_pool.ctx.erase(devices[g]); return rc; } }
            ctxs[g] = c;
        }
    }
    if (!records) cap = 0;
    // resolve the dealing mode once (on the first device) so that every device deals the same way
    {
        std::lock_guard<std::mutex> lk(ctxs[0]->mu);
        CUDA_TRY(cudaSetDevice(ctxs[0]->device));
        WaveCfg w;
        int rc = resolve_cfg(ctxs[0], cons, w);
        if (rc) return rc;
    }
    std::vector<int> rcs(G, BMC_OK);
    std::vector<std::string> errs(G);
    std::vector<bmc_stats> sts(G);
    std::vector<uint64_t> slot(G + 1, 0);
    const uint64_t cap_g = cap / (uint64_t)G;
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g)
        th.emplace_back([&, g] {
            bmc_ctx *ctx = ctxs[g];
            std::lock_guard<std::mutex> lk(ctx->mu);
            auto run = [&]() -> int {
                CUDA_TRY(cudaSetDevice(ctx->device));
                uint64_t first = first_index, count = 0, target = 0;
                if (n_accept_target) {
                    first = first_index + ((uint64_t)g << 40);
                    target = (n_accept_target + (uint64_t)G - 1) / (uint64_t)G;
                    count = n_raw ? (n_raw + (uint64_t)G - 1) / (uint64_t)G : 0;
                } else {
                    const uint64_t base = n_raw / (uint64_t)G, extra = n_raw % (uint64_t)G;
                    count = base + ((uint64_t)g < extra ? 1 : 0);
                    first = first_index + (uint64_t)g * base + ((uint64_t)g < extra ? (uint64_t)g : extra);
                }
                CUDA_TRY(cudaMemsetAsync(ctx->d_tallies, 0, sizeof(bmc_tallies), ctx->stream));
                if (!n_accept_target && count == 0) { sts[g] = bmc_stats{}; return cudaStreamSynchronize(ctx->stream) == cudaSuccess ? BMC_OK : BMC_E_CUDA; }
                int rc = ensure_scratch(ctx, cap_g * sizeof(bmc_record));
                if (rc) return rc;
                return sample_core(ctx, cons, seed, first, count, target, wave, tally_mask, cap_g ? ctx->d_rec : nullptr, cap_g,
                                   ctx->d_tallies, &sts[g], cap_g ? (void *)(records + (uint64_t)g * cap_g) : nullptr, 0u);
            };
            rcs[g] = run();
            if (rcs[g]) errs[g] = g_err;
        });
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return fail(rcs[g], "device " + std::to_string(devices[g]) + ": " + errs[g]);
    // ---- the one exchange: sum of the tally buffers ----
    bmc_tallies total;
    memset(&total, 0, sizeof total);
    bool p2p = G > 1;
    for (int g = 1; g < G && p2p; ++g) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctxs[0]->device, ctxs[g]->device) != cudaSuccess || !can) p2p = false;
    }
    if (p2p) {
        std::lock_guard<std::mutex> lk(ctxs[0]->mu);
        CUDA_TRY(cudaSetDevice(ctxs[0]->device));
        PeerTallies peers{};
        peers.n = G - 1;
        for (int g = 1; g < G; ++g) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[g]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { p2p = false; break; }
            cudaGetLastError();
            peers.p[g - 1] = ctxs[g]->d_tallies;
        }
        if (p2p) {
            reduce_peer_tallies_kernel<<<2, 256, 0, ctxs[0]->stream>>>(ctxs[0]->d_tallies, peers, (int)BMC_TALLY_WORDS);
            ctxs[0]->launches++;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(&total, ctxs[0]->d_tallies, sizeof total, cudaMemcpyDeviceToHost, ctxs[0]->stream));
            CUDA_TRY(cudaStreamSynchronize(ctxs[0]->stream));
        }
    }
    if (!p2p) {
        uint64_t *acc = reinterpret_cast<uint64_t *>(&total);
        memset(&total, 0, sizeof total);
        for (int g = 0; g < G; ++g) {
            bmc_tallies t;
            std::lock_guard<std::mutex> lk(ctxs[g]->mu);
            CUDA_TRY(cudaSetDevice(ctxs[g]->device));
            CUDA_TRY(cudaMemcpyAsync(&t, ctxs[g]->d_tallies, sizeof t, cudaMemcpyDeviceToHost, ctxs[g]->stream));
            CUDA_TRY(cudaStreamSynchronize(ctxs[g]->stream));
            const uint64_t *src = reinterpret_cast<const uint64_t *>(&t);
            for (size_t i = 0; i < BMC_TALLY_WORDS; ++i) acc[i] += src[i];
        }
    }
    // ---- close the gaps between the devices' record segments ----
    uint64_t stored = 0;
    for (int g = 0; g < G; ++g) {
        if (cap_g && g > 0 && sts[g].stored) memmove(records + stored, records + (uint64_t)g * cap_g, sts[g].stored * sizeof(bmc_record));
        stored += cap_g ? sts[g].stored : 0;
    }
    total.stored = stored;
    if (tallies) *tallies = total;
    if (stats) {
        bmc_stats s{};
        for (int g = 0; g < G; ++g) {
            s.raw += sts[g].raw; s.voided += sts[g].voided; s.accepted += sts[g].accepted;
            s.waves += sts[g].waves; s.launches += sts[g].launches;
            if (sts[g].kernel_ms > s.kernel_ms) s.kernel_ms = sts[g].kernel_ms;      // device time: the slowest device's
        }
        s.stored = stored;
        s.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *stats = s;
    }
    return BMC_OK;
}

// ---- device distgen (bridge.cl:1018-1045) --------------------------------------------------------------------
int bmc_distgen(bmc_ctx *ctx, uint64_t cards, const bmc_seat_spec *spec, uint16_t *patterns, uint64_t *weights, uint32_t cap,
                uint32_t *n_patterns, uint64_t *total_weight, uint64_t *hcp_pmf, uint64_t *len_pmf)
{
    if (!ctx || !n_patterns) return fail(BMC_E_INVALID, "bmc_distgen: NULL argument");
    *n_patterns = 0;
    if (cards & ~BMC_LANE_MASK) return fail(BMC_E_INVALID, "bmc_distgen: cards is not a lane mask");
    // the range spec as distgen reads it: :hcp (lo hi); suit (lo hi (shcp-lo shcp-hi)); an open upper length
    // bound is the number of LOW cards left in the suit (bridge.cl:1000-1001 `(or max supply)`)
    SeatRanges r;
    ClassSpec cs{};
    cs.hon_avail = 0;
    for (int s = 0; s < 4; ++s) {
        const uint32_t lane = (uint32_t)(cards >> (16 * s)) & 0x1FFFu;
        cs.n_low[s] = (uint32_t)__builtin_popcount(lane & 0x1FFu);
        for (int h = 0; h < 4; ++h)
            if ((lane >> (9 + h)) & 1u) cs.hon_avail |= 1u << (15 - (4 * s + h));
    }
    cs.hand_size = 13;
    if (spec) {
        if (spec->hcp[0] >= 0) r.lo[F_HCP] = spec->hcp[0];
        if (spec->hcp[1] >= 0) r.hi[F_HCP] = spec->hcp[1];
        for (int s = 0; s < 4; ++s) {
            if (spec->len[s][0] >= 0) r.lo[F_C + s] = spec->len[s][0];
            r.hi[F_C + s] = spec->len[s][1] >= 0 ? spec->len[s][1] : (int)cs.n_low[s];
            if (spec->suit_hcp[s][0] >= 0) r.lo[F_CP + s] = spec->suit_hcp[s][0];
            if (spec->suit_hcp[s][1] >= 0) r.hi[F_CP + s] = spec->suit_hcp[s][1];
        }
    } else
        for (int s = 0; s < 4; ++s) r.hi[F_C + s] = (int)cs.n_low[s];
    for (int s = 0; s < 4; ++s) { if (r.hi[F_C + s] > 13) r.hi[F_C + s] = 13; if (r.hi[F_CP + s] > 10) r.hi[F_CP + s] = 10; }
    if (r.hi[F_HCP] > 40) r.hi[F_HCP] = 40;
    SeatC sc{};
    const bool empty = r.empty();                       // an empty range admits no pattern at all
    if (!empty) ranges_to_seatc(r, sc);
    sc.prog = 0xFFFFFFFFu; sc.active = 1;
    cs.seat = sc;
    cs.prog = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    PatternScan ps;
    if (empty) { ps.w.assign(DG_PATTERNS, 0); ps.cnt.assign(DG_PATTERNS, 0); ps.flag.assign(DG_PATTERNS, 0); memset(ps.hcp_pmf, 0, sizeof ps.hcp_pmf); memset(ps.len_pmf, 0, sizeof ps.len_pmf); }
    else { int rc = distgen_count(ctx, cs, ps, hcp_pmf || len_pmf); if (rc) return rc; }
    uint32_t n = 0;
    uint64_t total = 0;
    // rows in the generator's order: binary counting, club jack most significant = ascending pattern value.
    // bridge.cl:955-957: a deck without any honour has no pattern at all (hc-permuts returns nil).
    for (uint32_t p = 0; p < DG_PATTERNS && cs.hon_avail; ++p) {
        if (!ps.flag[p]) continue;
        if (n < cap) { if (patterns) patterns[n] = (uint16_t)p; if (weights) weights[n] = ps.w[p]; }
        total += ps.w[p];
        ++n;
    }
    *n_patterns = n;
    if (total_weight) *total_weight = total;
    if (hcp_pmf) for (int i = 0; i < 38; ++i) hcp_pmf[i] = empty ? 0 : ps.hcp_pmf[i];
    if (len_pmf) for (int i = 0; i < 56; ++i) len_pmf[i] = empty ? 0 : ps.len_pmf[i];
    if (n > cap && (patterns || weights)) return fail(BMC_E_CAPACITY, "bmc_distgen: " + std::to_string(n) + " patterns, capacity " + std::to_string(cap));
    return BMC_OK;
}

int bmc_constraints_primary_acceptance(const bmc_constraints *c, uint64_t *hands, uint64_t *classes)
{
    if (!c || !hands) return fail(BMC_E_INVALID, "bmc_constraints_primary_acceptance: NULL argument");
    if (!c->x_known) return fail(BMC_E_INVALID, "bmc_constraints_primary_acceptance: not enumerated yet (first sampling call in mode 0 or 4)");
    *hands = c->x_weight;
    if (classes) *classes = c->x_classes;
    return BMC_OK;
}

// ---- filter ---------------------------------------------------------------------
// One pass of K2 over device buffers.  A rule table (scheme) is interpreted row by row, or -- scheme->jit 1, or 2
// (default) for large (>= 2^21 deals) or repeated jobs -- evaluated by the NVRTC-compiled form of its rows.
static int filter_core(bmc_ctx *ctx, const bmc_constraints *cons, const bmc_scheme *scheme, const uint4 *d_rec, uint64_t k,
                       uint8_t *d_acc, bmc_features *d_feat)
{
    ConsDev dev{};
    const bool has_cons = cons && cons->has_cons;
    if (has_cons) { int rc = cons_upload(ctx, cons, dev); if (rc) return rc; }
    SchemeDev sd;
    SchemeTables *st = nullptr;
    { int rc = scheme_upload(ctx, scheme, sd, &st); if (rc) return rc; }
    const JitKernel *jit = nullptr;
    bmc_scheme *sch = const_cast<bmc_scheme *>(scheme);
    if (st && sd.n_rows) {
        std::lock_guard<std::mutex> lk(sch->mu);
        const bool want = sch->jit == 1 || (sch->jit == 2 && (k >= (1ull << 21) || sch->interp_ms > 500.0) && JitApi::get().ok);
        if (want) {
            int rc = scheme_jit_build(ctx, sch, st);
            if (rc) {
                if (sch->jit == 1) return rc;
                sch->jit = 0;                        // auto: the build failed once -- keep interpreting
            } else jit = &st->jit;
        }
    }
    const unsigned blocks = (unsigned)((k + 255) / 256);
    int hc = has_cons ? 1 : 0;
    unsigned long long kk = k;
    FeatOut *fo = reinterpret_cast<FeatOut *>(d_feat);
    const bool timed = st && sd.n_rows && !jit;
    if (timed) CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    if (jit) {
        void *params[] = {&dev, &hc, &sd, (void *)&d_rec, &kk, &d_acc, &fo};
        const CUresult cr = JitApi::get().LaunchKernel(jit->func, blocks, 1, 1, 256, 1, 1, 0, (CUstream)ctx->stream, params, nullptr);
        if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "cuLaunchKernel (jit filter) failed: " + std::to_string((int)cr));
    } else filter_kernel<<<blocks, 256, 0, ctx->stream>>>(dev, hc, sd, d_rec, kk, d_acc, fo);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    if (timed) {
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        std::lock_guard<std::mutex> lk(sch->mu);
        sch->interp_ms += ms;
    }
    return BMC_OK;
}

int bmc_filter(bmc_ctx *ctx, const bmc_constraints *cons, const bmc_scheme *scheme, const bmc_record *records, uint64_t n,
               uint8_t *accept, bmc_features *feats)
{
    if (!ctx) return fail(BMC_E_INVALID, "ctx is NULL");
    if (n == 0) return BMC_OK;
    if (!records) return fail(BMC_E_INVALID, "records is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t chunk = 1ull << 22;       // bounded device footprint: 4 Mi deals per pass
    uint4 *d_rec = nullptr;
    uint8_t *d_acc = nullptr;
    bmc_features *d_feat = nullptr;
    const uint64_t m = n < chunk ? n : chunk;
    CUDA_TRY(cudaMalloc(&d_rec, m * sizeof(uint4)));
    cudaError_t e = cudaSuccess;
    if (accept) e = cudaMalloc(&d_acc, m);
    if (e == cudaSuccess && feats) e = cudaMalloc(&d_feat, m * sizeof(bmc_features));
    int rc = BMC_OK;
    for (uint64_t off = 0; off < n && e == cudaSuccess && rc == BMC_OK; off += chunk) {
        const uint64_t k = n - off < chunk ? n - off : chunk;
        e = cudaMemcpyAsync(d_rec, records + off, k * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) break;
        rc = filter_core(ctx, cons, scheme, d_rec, k, d_acc, d_feat);
        if (rc) break;
        if (accept) e = cudaMemcpyAsync(accept + off, d_acc, k, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && feats) e = cudaMemcpyAsync(feats + off, d_feat, k * sizeof(bmc_features), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (rc == BMC_OK && e != cudaSuccess) rc = fail(BMC_E_CUDA, std::string("bmc_filter: ") + cudaGetErrorString(e));
    cudaFree(d_rec); cudaFree(d_acc); cudaFree(d_feat);
    return rc;
}

// device-buffer form: records_dev / accept_dev / feats_dev are device pointers (either output may be NULL); asynchronous
// on the context's stream like bmc_sample_dev's outputs
int bmc_filter_dev(bmc_ctx *ctx, const bmc_constraints *cons, const bmc_scheme *scheme, const void *records_dev, uint64_t n,
                   void *accept_dev, void *feats_dev)
{
    if (!ctx) return fail(BMC_E_INVALID, "ctx is NULL");
    if (n == 0) return BMC_OK;
    if (!records_dev) return fail(BMC_E_INVALID, "records_dev is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return filter_core(ctx, cons, scheme, (const uint4 *)records_dev, n, (uint8_t *)accept_dev, (bmc_features *)feats_dev);
}

// ---- auction driver (bidding.cl:457-480) -------------------------------------------------------------------------
static int auction_core(bmc_ctx *ctx, const bmc_scheme *openings, const bmc_scheme *const *responses, int dealer, const uint4 *d_rec,
                        uint64_t k, bmc_auction_row *d_out)
{
    AuctionDev a{};
    a.dealer = dealer & 3;
    { int rc = scheme_upload(ctx, openings, a.open); if (rc) return rc; }
    for (uint32_t r = 0; r < a.open.n_rows; ++r) {
        int rc = scheme_upload(ctx, responses ? responses[r] : nullptr, a.resp[r]);
        if (rc) return rc;
    }
    auction_kernel<<<(unsigned)((k + 255) / 256), 256, 0, ctx->stream>>>(a, d_rec, k, reinterpret_cast<AuctionOut *>(d_out));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return BMC_OK;
}

int bmc_auction(bmc_ctx *ctx, const bmc_scheme *openings, const bmc_scheme *const *responses, int dealer, const bmc_record *records,
                uint64_t n, bmc_auction_row *out)
{
    if (!ctx || !openings || !out) return fail(BMC_E_INVALID, "bmc_auction: NULL argument");
    if (openings->row_off.size() > (size_t)AUCTION_MAX_OPENINGS) return fail(BMC_E_CAPACITY, "bmc_auction: at most 32 opening rows");
    if (dealer < 0 || dealer > 3) return fail(BMC_E_INVALID, "bmc_auction: dealer is a seat 0..3");
    if (n == 0) return BMC_OK;
    if (!records) return fail(BMC_E_INVALID, "records is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t chunk = 1ull << 22;
    const uint64_t m = n < chunk ? n : chunk;
    { int rc = ensure_scratch(ctx, m * (sizeof(bmc_record) + sizeof(bmc_auction_row))); if (rc) return rc; }
    uint4 *d_rec = (uint4 *)ctx->d_rec;
    bmc_auction_row *d_out = (bmc_auction_row *)((char *)ctx->d_rec + m * sizeof(bmc_record));
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t k = n - off < chunk ? n - off : chunk;
        CUDA_TRY(cudaMemcpyAsync(d_rec, records + off, k * sizeof(bmc_record), cudaMemcpyHostToDevice, ctx->stream));
        int rc = auction_core(ctx, openings, responses, dealer, d_rec, k, d_out);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(out + off, d_out, k * sizeof(bmc_auction_row), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return BMC_OK;
}

int bmc_auction_dev(bmc_ctx *ctx, const bmc_scheme *openings, const bmc_scheme *const *responses, int dealer, const void *records_dev,
                    uint64_t n, void *out_dev)
{
    if (!ctx || !openings || !out_dev) return fail(BMC_E_INVALID, "bmc_auction_dev: NULL argument");
    if (openings->row_off.size() > (size_t)AUCTION_MAX_OPENINGS) return fail(BMC_E_CAPACITY, "bmc_auction_dev: at most 32 opening rows");
    if (dealer < 0 || dealer > 3) return fail(BMC_E_INVALID, "bmc_auction_dev: dealer is a seat 0..3");
    if (n == 0) return BMC_OK;
    if (!records_dev) return fail(BMC_E_INVALID, "records_dev is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return auction_core(ctx, openings, responses, dealer, (const uint4 *)records_dev, n, (bmc_auction_row *)out_dev);
}

// ---- scoring --------------------------------------------------------------------
int bmc_score(bmc_ctx *ctx, const bmc_contract *contracts, const uint8_t *vulnerable, uint32_t nc, const uint8_t *tricks,
              uint64_t n, int32_t *scores)
{
    if (!ctx || !contracts || !tricks || !scores || nc == 0) return fail(BMC_E_INVALID, "bmc_score: NULL argument");
    if (n == 0) return BMC_OK;
    { int rc = check_contracts(contracts, nc); if (rc) return rc; }
    const uint64_t total = n * nc;
    for (uint64_t i = 0; i < total; ++i) if (tricks[i] > 13) return fail(BMC_E_INVALID, "tricks out of 0..13");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    bmc_contract *d_c = nullptr; uint8_t *d_v = nullptr, *d_t = nullptr; int32_t *d_s = nullptr;
    cudaError_t e = cudaMalloc(&d_c, nc * sizeof(bmc_contract));
    if (e == cudaSuccess) e = cudaMalloc(&d_v, nc);
    if (e == cudaSuccess) e = cudaMalloc(&d_t, total);
    if (e == cudaSuccess) e = cudaMalloc(&d_s, total * sizeof(int32_t));
    std::vector<uint8_t> vul(nc, 0);
    if (vulnerable) memcpy(vul.data(), vulnerable, nc);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_c, contracts, nc * sizeof(bmc_contract), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_v, vul.data(), nc, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_t, tricks, total, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const uint64_t want = (total / 4 + 255) / 256 + 1;
        const unsigned blocks = (unsigned)(want < (uint64_t)ctx->sms * 16 ? want : (uint64_t)ctx->sms * 16);
        score_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const ContractDev *>(d_c), d_v, nc, d_t, total, d_s);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(scores, d_s, total * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_c); cudaFree(d_v); cudaFree(d_t); cudaFree(d_s);
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("bmc_score: ") + cudaGetErrorString(e));
    return BMC_OK;
}

int bmc_match_points(bmc_ctx *ctx, const int32_t *a, const int32_t *b, uint64_t n, int8_t *imps, uint8_t *status)
{
    if (!ctx || !a || !b || !imps) return fail(BMC_E_INVALID, "bmc_match_points: NULL argument");
    if (n == 0) return BMC_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    int32_t *d_a = nullptr, *d_b = nullptr; int8_t *d_i = nullptr; uint8_t *d_s = nullptr;
    cudaError_t e = cudaMalloc(&d_a, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_b, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_i, n);
    if (e == cudaSuccess) e = cudaMalloc(&d_s, n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_a, a, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_b, b, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((n + 255) / 256 < (uint64_t)ctx->sms * 16 ? (n + 255) / 256 : (uint64_t)ctx->sms * 16);
        imps_kernel<<<blocks, 256, 0, ctx->stream>>>(d_a, d_b, n, d_i, d_s);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(imps, d_i, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && status) e = cudaMemcpyAsync(status, d_s, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_i); cudaFree(d_s);
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("bmc_match_points: ") + cudaGetErrorString(e));
    return BMC_OK;
}

int bmc_score_tally(bmc_ctx *ctx, const bmc_contract *contracts, const uint8_t *vulnerable, uint32_t nc, const uint8_t *pairs,
                    uint32_t n_pairs, uint64_t seed, uint64_t first_index, uint64_t n, uint64_t *tricks_hist, uint64_t *imp_hist,
                    uint64_t *imp_dropped, int32_t *score_table)
{
    if (!ctx || !contracts || nc == 0 || nc > 8 || n_pairs > 8 || !tricks_hist) return fail(BMC_E_INVALID, "bmc_score_tally: bad argument");
    if (n_pairs && (!pairs || !imp_hist)) return fail(BMC_E_INVALID, "bmc_score_tally: pairs / imp_hist is NULL");
    ContractTable t{};
    { int rc = fill_contract_table(t, contracts, vulnerable, nc, pairs, n_pairs); if (rc) return rc; }
    if (score_table)
        for (uint32_t c = 0; c < nc; ++c)
            for (int tr = 0; tr < 14; ++tr) score_table[c * 14 + tr] = base_score_dev(t.c[c], tr, t.vul[c]);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, SCORE_WORDS * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, SCORE_WORDS * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess && n) {
        uint64_t want = (n + 255) / 256;
        const unsigned blocks = (unsigned)(want < (uint64_t)ctx->sms * 8 ? want : (uint64_t)ctx->sms * 8);
        score_tally_kernel<<<blocks, 256, 0, ctx->stream>>>(t, philox_key(seed), first_index, n, d);
        ctx->launches++;
        e = cudaGetLastError();
    }
    std::vector<unsigned long long> h(SCORE_WORDS);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("bmc_score_tally: ") + cudaGetErrorString(e));
    for (uint32_t c = 0; c < nc; ++c)
        for (int tr = 0; tr < 14; ++tr) tricks_hist[c * 14 + tr] = h[c * 14 + tr];
    for (uint32_t p = 0; p < n_pairs; ++p) {
        for (int v = 0; v < 49; ++v) imp_hist[p * 49 + v] = h[SCORE_TR_WORDS + p * 49 + v];
        if (imp_dropped) imp_dropped[p] = h[SCORE_TR_WORDS + SCORE_IMP_WORDS + p];
    }
    return BMC_OK;
}

// ---- hist-base over measure tuples (bridge.cl:1325-1330) ---------------------------------------------
__global__ void gather_first_kernel(const uint32_t *__restrict__ idx_sorted, const unsigned long long *__restrict__ offsets,
                                    uint32_t runs, uint32_t *__restrict__ first)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < runs) first[i] = idx_sorted[offsets[i]];
}

int bmc_hist_base(bmc_ctx *ctx, const bmc_record *records, uint64_t n, const bmc_measure *measures, uint32_t k,
                  uint8_t *tuples, uint64_t *counts, uint64_t cap, uint64_t *n_unique)
{
    if (!ctx || !measures || k == 0 || k > 8 || !n_unique) return fail(BMC_E_INVALID, "bmc_hist_base: bad argument");
    *n_unique = 0;
    if (n == 0) return BMC_OK;
    if (!records) return fail(BMC_E_INVALID, "bmc_hist_base: records is NULL");
    if (n >= (1ull << 31)) return fail(BMC_E_CAPACITY, "bmc_hist_base: at most 2^31 - 1 deals per call");
    MeasureDev md{};
    md.k = k;
    for (uint32_t j = 0; j < k; ++j) {
        if (measures[j].kind > M_LONGEST || measures[j].seat > 3 || measures[j].suit > 3) return fail(BMC_E_INVALID, "bmc_hist_base: bad measure");
        md.kind[j] = measures[j].kind; md.seat[j] = measures[j].seat; md.suit[j] = measures[j].suit;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint4 *d_rec = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr, *d_uniq = nullptr, *d_off = nullptr;
    uint32_t *d_idx = nullptr, *d_idx2 = nullptr, *d_cnt = nullptr, *d_runs = nullptr, *d_first = nullptr;
    uint8_t *d_lut = nullptr;
    void *d_tmp = nullptr;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    alloc((void **)&d_rec, n * sizeof(uint4));
    alloc((void **)&d_keys, n * 8); alloc((void **)&d_keys2, n * 8); alloc((void **)&d_uniq, n * 8);
    alloc((void **)&d_idx, n * 4); alloc((void **)&d_idx2, n * 4); alloc((void **)&d_cnt, n * 4); alloc((void **)&d_runs, 4);
    alloc((void **)&d_lut, sizeof(ShapeLut));
    std::vector<unsigned long long> h_uniq;
    std::vector<uint32_t> h_cnt, h_first;
    uint32_t runs = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rec, records, n * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);
    static const ShapeLut h_lut = make_shape_lut();
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_lut, h_lut.v, sizeof(ShapeLut), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        measure_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(md, d_rec, n, d_lut, d_keys, d_idx);
        ctx->launches++;
        e = cudaGetLastError();
    }
    size_t tmp_bytes = 0, tmp2 = 0;
    const int begin_bit = 8 * (8 - (int)k);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)n, begin_bit, 64, ctx->stream);
    if (e == cudaSuccess) e = cub::DeviceRunLengthEncode::Encode(nullptr, tmp2, d_keys2, d_uniq, d_cnt, d_runs, (int)n, ctx->stream);
    if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    alloc(&d_tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)n, begin_bit, 64, ctx->stream);
    if (e == cudaSuccess) e = cub::DeviceRunLengthEncode::Encode(d_tmp, tmp_bytes, d_keys2, d_uniq, d_cnt, d_runs, (int)n, ctx->stream);
    ctx->launches += 2;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&runs, d_runs, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && runs) {
        h_uniq.resize(runs); h_cnt.resize(runs); h_first.resize(runs);
        e = cudaMemcpyAsync(h_uniq.data(), d_uniq, runs * 8ull, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt.data(), d_cnt, runs * 4ull, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        std::vector<unsigned long long> off(runs);
        unsigned long long acc = 0;
        for (uint32_t i = 0; i < runs; ++i) { off[i] = acc; acc += h_cnt[i]; }
        alloc((void **)&d_off, runs * 8ull); alloc((void **)&d_first, runs * 4ull);
        // same stream as the kernel that reads it (ctx->stream is non-blocking: the legacy stream would not order them)
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, off.data(), runs * 8ull, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            gather_first_kernel<<<(runs + 255) / 256, 256, 0, ctx->stream>>>(d_idx2, d_off, runs, d_first);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_first.data(), d_first, runs * 4ull, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(d_rec); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_uniq); cudaFree(d_off); cudaFree(d_idx); cudaFree(d_idx2);
    cudaFree(d_cnt); cudaFree(d_runs); cudaFree(d_first); cudaFree(d_lut); cudaFree(d_tmp);
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("bmc_hist_base: ") + cudaGetErrorString(e));
    *n_unique = runs;
    if (runs > cap) return fail(BMC_E_CAPACITY, "bmc_hist_base: " + std::to_string(runs) + " distinct tuples, capacity " + std::to_string(cap));
    if (runs && (!tuples || !counts)) return fail(BMC_E_INVALID, "bmc_hist_base: output buffers are NULL");
    // first-seen order, like hist-base
    std::vector<uint32_t> order(runs);
    for (uint32_t i = 0; i < runs; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h_first[a] < h_first[b]; });
    for (uint32_t o = 0; o < runs; ++o) {
        const uint32_t i = order[o];
        for (uint32_t j = 0; j < k; ++j) tuples[(uint64_t)o * k + j] = (uint8_t)(h_uniq[i] >> (8 * (7 - j)));
        counts[o] = h_cnt[i];
    }
    return BMC_OK;
}

// ---- measurement ------------------------------------------------------------------
// ---------------------------------------------------------------------------
// play-deal (bridge.cl:2961): one warp per deal (playdeal.cuh)
// ---------------------------------------------------------------------------
// the kernels live in their own translation unit (playdeal_kernels.cu): they are compiled for code size
static int pd_blocks_per_sm()
{
    const char *e = getenv("BMC_PD_BLOCKS");
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 16 ? 16 : v);
}

static int play_deal_launch(bmc_ctx *ctx, const void *d_records, const void *d_hands4, uint64_t n, const int8_t *d_trump,
                            const uint8_t *d_declarer, int per_deal, uint32_t arena_kb, void *d_out)
{
    // arena_kb = 0: 1 MB per resident deal in the first pass (2.4 GB of HBM for 148 x 16 deals; the largest of 3 000
    // random deals needs 0.5 MB), and the deals that outgrow it are played again with 8 MB each -- a second, nearly empty
    // launch, which is why the first pass is sized to make it rare; an explicit arena_kb is used as it is, without a retry
    const bool adaptive = arena_kb == 0;
    const char *env_kb = getenv("BMC_PD_ARENA_KB");
    const uint32_t first_kb = env_kb && atoi(env_kb) >= 16 ? (uint32_t)atoi(env_kb) : 1024u;
    const int timing = getenv("BMC_PD_TIMING") != nullptr;
    const uint64_t arena = (uint64_t)(adaptive ? first_kb : arena_kb) * 1024u;
    if (arena > (1ull << 31)) return fail(BMC_E_INVALID, "bmc_play_deal: arena_kb too large");
    if (n > 0xFFFFFFFFull) return fail(BMC_E_INVALID, "bmc_play_deal: more than 2^32 - 1 deals in one call");
    constexpr int PD_WARPS = bmc_pd::WARPS;
    int per_sm = 1;
    CUDA_TRY(bmc_pd::configure(&per_sm));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > pd_blocks_per_sm()) per_sm = pd_blocks_per_sm();
    uint64_t blocks = (n + PD_WARPS - 1) / PD_WARPS;
    const uint64_t max_blocks = (uint64_t)ctx->sms * per_sm;
    if (blocks > max_blocks) blocks = max_blocks;
    unsigned long long *d_next = nullptr;
    CUDA_TRY(cudaMalloc(&d_next, 3 * sizeof(unsigned long long)));
    struct Free { unsigned long long *p; ~Free() { cudaFree(p); } } free_next{d_next};
    CUDA_TRY(cudaMemsetAsync(d_next, 0, 3 * sizeof(unsigned long long), ctx->stream));
    // the arenas (one per resident warp): grow-only, halving the number of concurrent deals if the device cannot hold them
    for (;;) {
        const uint64_t need = blocks * PD_WARPS * arena;
        if (need <= ctx->arena_bytes) break;
        if (ctx->d_arena) cudaFree(ctx->d_arena);
        ctx->d_arena = nullptr; ctx->arena_bytes = 0;
        if (cudaMalloc(&ctx->d_arena, need) == cudaSuccess) { ctx->arena_bytes = need; break; }
        cudaGetLastError();
        if (blocks <= 1) return fail(BMC_E_NOMEM, "bmc_play_deal: cannot allocate the per-deal arenas");
        blocks = (blocks + 1) / 2;
    }
    // A batch ends with its longest deal, and a deal takes anything from a hundredth to fifteen times the mean.  When the
    // batch is more than a few rounds of the resident warps, a probe pass (the set-up of the two strategies, 5 % of the
    // work) scores every deal, and the deals are played longest-expected first: 1.3 x on 32 768 deals.
    unsigned int *d_order = nullptr;
    struct FreeU { unsigned int *&p; ~FreeU() { if (p) cudaFree(p); } } free_order{d_order};
    const char *env_order = getenv("BMC_PD_ORDER");
    const bool ordered = env_order ? atoi(env_order) != 0 : n >= 2ull * blocks * PD_WARPS;
    if (ordered) {
        unsigned int *d_score = nullptr;
        CUDA_TRY(cudaMalloc(&d_score, n * sizeof(unsigned int)));
        struct FreeS { unsigned int *p; ~FreeS() { cudaFree(p); } } free_score{d_score};
        bmc_pd::launch((unsigned)blocks, ctx->stream, d_records, d_hands4, nullptr, n, d_trump, d_declarer, per_deal, ctx->d_arena,
                       (uint32_t)arena, d_out, d_next + 2, 0, d_score);
        ctx->launches++;
        std::vector<unsigned int> score(n), order(n);
        CUDA_TRY(cudaMemcpyAsync(score.data(), d_score, n * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        for (uint64_t i = 0; i < n; ++i) order[i] = (unsigned int)i;
        std::stable_sort(order.begin(), order.end(), [&](unsigned int a, unsigned int b) { return score[a] > score[b]; });
        CUDA_TRY(cudaMalloc(&d_order, n * sizeof(unsigned int)));
        CUDA_TRY(cudaMemcpyAsync(d_order, order.data(), n * sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));            // `order` leaves scope
    }
    bmc_pd::launch((unsigned)blocks, ctx->stream, d_records, d_hands4, d_order, n, d_trump, d_declarer, per_deal, ctx->d_arena,
                   (uint32_t)arena, d_out, d_next, timing);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && adaptive) {
        unsigned int *d_idx = nullptr, *d_cnt = nullptr, h_cnt = 0;
        e = cudaMalloc(&d_idx, n * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMalloc(&d_cnt, sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int), ctx->stream);
        if (e == cudaSuccess) {
            bmc_pd::launch_failed(ctx->stream, d_out, n, d_idx, d_cnt);
            ctx->launches++;
            e = cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess && h_cnt) {
            const uint64_t big = 8ull << 20;
            uint64_t b2 = ctx->arena_bytes / (big * PD_WARPS);
            if (b2 == 0) {                                       // the first pass used less than one block of 8 MB arenas: grow
                cudaFree(ctx->d_arena);
                ctx->d_arena = nullptr; ctx->arena_bytes = 0;
                if (cudaMalloc(&ctx->d_arena, PD_WARPS * big) != cudaSuccess) { cudaGetLastError(); cudaFree(d_idx); cudaFree(d_cnt); return fail(BMC_E_NOMEM, "bmc_play_deal: cannot allocate the retry arenas"); }
                ctx->arena_bytes = PD_WARPS * big;
                b2 = 1;
            }
            const uint64_t want = ((uint64_t)h_cnt + PD_WARPS - 1) / PD_WARPS;
            if (b2 > want) b2 = want;
            if (b2 > max_blocks) b2 = max_blocks;
            bmc_pd::launch((unsigned)b2, ctx->stream, d_records, d_hands4, d_idx, h_cnt, d_trump, d_declarer, per_deal, ctx->d_arena,
                           (uint32_t)big, d_out, d_next + 1, timing);
            ctx->launches++;
            e = cudaGetLastError();
        }
        cudaFree(d_idx); cudaFree(d_cnt);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("bmc_play_deal: ") + cudaGetErrorString(e));
    return BMC_OK;
}

int bmc_play_deal_dev(bmc_ctx *ctx, const void *d_records, uint64_t n, const int8_t *d_trump, const uint8_t *d_declarer,
                      int per_deal, uint32_t arena_kb, void *d_out)
{
    if (!ctx || !d_records || !d_trump || !d_declarer || !d_out) return fail(BMC_E_INVALID, "bmc_play_deal_dev: NULL argument");
    if (n == 0) return BMC_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    return play_deal_launch(ctx, d_records, nullptr, n, d_trump, d_declarer, per_deal, arena_kb, d_out);
}

static int play_deal_host(bmc_ctx *ctx, const void *records, const void *hands4, uint64_t n, const int8_t *trump,
                          const uint8_t *declarer, int per_deal, uint32_t arena_kb, bmc_play_result *out)
{
    const uint64_t np = per_deal ? n : 1;
    for (uint64_t i = 0; i < np; ++i)
        if (trump[i] < -1 || trump[i] > 3 || (declarer && declarer[i] > 3)) return fail(BMC_E_INVALID, "play-deal: trump in -1..3, declarer in 0..3");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t in_bytes = records ? n * sizeof(bmc_record) : n * 4 * sizeof(uint64_t);
    void *d_in = nullptr, *d_o = nullptr; int8_t *d_t = nullptr; uint8_t *d_d = nullptr;
    cudaError_t e = cudaMalloc(&d_in, in_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, n * sizeof(bmc_play_result));
    if (e == cudaSuccess) e = cudaMalloc(&d_t, np);
    if (e == cudaSuccess) e = cudaMalloc(&d_d, np);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, records ? records : hands4, in_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_t, trump, np, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && declarer) e = cudaMemcpyAsync(d_d, declarer, np, cudaMemcpyHostToDevice, ctx->stream);
    int rc = BMC_OK;
    if (e == cudaSuccess) rc = play_deal_launch(ctx, records ? d_in : nullptr, records ? nullptr : d_in, n, d_t, d_d, per_deal, arena_kb, d_o);
    if (e == cudaSuccess && rc == BMC_OK) e = cudaMemcpyAsync(out, d_o, n * sizeof(bmc_play_result), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == BMC_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in); cudaFree(d_o); cudaFree(d_t); cudaFree(d_d);
    if (rc != BMC_OK) return rc;
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("play-deal: ") + cudaGetErrorString(e));
    return BMC_OK;
}

int bmc_play_deal(bmc_ctx *ctx, const bmc_record *records, uint64_t n, const int8_t *trump, const uint8_t *declarer,
                  int per_deal, uint32_t arena_kb, bmc_play_result *out)
{
    if (!ctx || !records || !trump || !declarer || !out) return fail(BMC_E_INVALID, "bmc_play_deal: NULL argument");
    if (n == 0) return BMC_OK;
    return play_deal_host(ctx, records, nullptr, n, trump, declarer, per_deal, arena_kb, out);
}

int bmc_play_hands(bmc_ctx *ctx, const uint64_t *hands, uint64_t n, const int8_t *trump, int per_deal, uint32_t arena_kb,
                   bmc_play_result *out)
{
    if (!ctx || !hands || !trump || !out) return fail(BMC_E_INVALID, "bmc_play_hands: NULL argument");
    if (n == 0) return BMC_OK;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t seen = 0;
        for (int k = 0; k < 4; ++k) {
            const uint64_t m = hands[4 * i + k];
            if ((m & ~BMC_LANE_MASK) || (m & seen)) return fail(BMC_E_INVALID, "bmc_play_hands: hands overlap or use bits outside the lane mask");
            seen |= m;
        }
    }
    return play_deal_host(ctx, nullptr, hands, n, trump, nullptr, per_deal, arena_kb, out);
}

int bmc_best_tricks(bmc_ctx *ctx, const bmc_record *records, uint64_t n, const int8_t *trump, const uint8_t *declarer,
                    int per_deal, uint32_t arena_kb, uint8_t *tricks)
{
    if (!tricks) return fail(BMC_E_INVALID, "bmc_best_tricks: NULL argument");
    std::vector<bmc_play_result> res(n);
    const int rc = bmc_play_deal(ctx, records, n, trump, declarer, per_deal, arena_kb, res.data());
    if (rc != BMC_OK) return rc;
    for (uint64_t i = 0; i < n; ++i) tricks[i] = res[i].status == 0 ? res[i].score[1] : 255;
    return BMC_OK;
}

int bmc_int_peak(bmc_ctx *ctx, double *ops_per_s, float *ms_out)
{
    if (!ctx || !ops_per_s) return fail(BMC_E_INVALID, "bmc_int_peak: NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int blocks = ctx->sms * 8;
    uint32_t *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)blocks * 256 * sizeof(uint32_t)));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        int_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(12345u + rep, d);
        ctx->launches++;
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(d); return fail(BMC_E_CUDA, std::string("int_peak: ") + cudaGetErrorString(e)); }
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d);
    const double ops = (double)blocks * 256.0 * PEAK_ITERS * PEAK_OPS_PER_ITER;
    *ops_per_s = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return BMC_OK;
}

}  // extern "C"
