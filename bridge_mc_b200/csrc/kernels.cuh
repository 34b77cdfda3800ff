// kernels.cuh -- the sampling kernels (device code only).
//
// Compiled twice: ahead of time by nvcc as part of bridgemc.cu (rule forms are interpreted from
// bytecode), and -- optionally, per constraint -- at run time by NVRTC with BMC_JIT defined, where
// `bmc_jit_seat_rule` is C++ generated from the constraint's rule forms (jit.hpp), the way the
// reference `eval`s (i.e. compiles) its rule forms (bidding.cl:60-74).
#pragma once
#include "deal.cuh"
#include "features.cuh"
#include "seatfirst.cuh"

#ifndef BMC_TALLY_HCP          /* NVRTC build: the few constants of bridgemc.h the kernels need */
#define BMC_TALLY_HCP   1u
#define BMC_TALLY_LEN   2u
#define BMC_TALLY_SHAPE 4u
#define BMC_LEN_BINS   14
#endif

using namespace bmc;

// ---------------------------------------------------------------------------
// device-side constraint / tally layout
// ---------------------------------------------------------------------------
struct ConsDev {
    SeatC seat[4];
    uint32_t n_active;       // number of constrained seats
    uint32_t primary;        // seat tested in stage A
    uint32_t n_prog;         // constrained seats with a rule program (seat-first mode: 0 = no second stage B)
    uint32_t b1_mask;        // compiled-rule (NVRTC) build of the seat-first kernel: seats fully tested in stage B1 (the
    uint32_t pad;            // primary and the most selective other seat, chosen by calibration); the rest follow in B2
    uint32_t sf_bmul;        // seat-first stage A2: bucket of a survivor = umulhi(u_hi, sf_bmul) (seatfirst.cuh)
    const uint32_t *code;    // device pointer, programs of all seats
    const uint4 *sf_tab;     // seat-first stage A1: class table of the primary seat (seatfirst.cuh), in the code buffer
    unsigned long long sf_theta;   // u < sf_theta: the honour holding has an admissible HCP
    unsigned long long sf_void;    // u >= sf_void = S * C(52,13) voids the index (probability 1.2e-8)
};

// shared-memory histogram (per warp): hcp[4][40] | len[4][4][16] | shape[4][40]
constexpr int SH_HCP = 0, SH_LEN = 160, SH_SHAPE = 160 + 256, SH_WORDS = 160 + 256 + 160;
constexpr int WARPS_PER_BLOCK = 8, THREADS = WARPS_PER_BLOCK * 32;
constexpr int QCAP = 64;     // per-warp queue of stage-A survivors

__constant__ uint8_t c_shape_lut[14 * 14 * 14];   // (len_c, len_d, len_h) -> sorted-shape class

struct SampleArgs {
    ConsDev cons;
    PhiloxKey key;
    FixDev fix;
    uint64_t first, n;
    uint4 *records;          // may be null
    uint64_t cap;
    unsigned long long *tallies;   // bmc_tallies as u64 words
    uint32_t tally_mask;
    uint32_t iters;          // warp-iterations per warp (seat-first: each covers 4 SF_CALLS indices per lane)
    uint32_t sf_q0;          // seat-first mode: initial Q of stage B (capacity 0 for the primary seat)
    uint32_t pad;
};

__device__ __forceinline__ void seat_words(const Planes &pl, int k, uint32_t &m0, uint32_t &m1)
{
    const uint32_t xl = (k & 1) ? 0u : 0xFFFFFFFFu, xh = (k & 2) ? 0u : 0xFFFFFFFFu;
    m0 = (pl.lo0 ^ xl) & (pl.hi0 ^ xh) & 0x1FFF1FFFu;
    m1 = (pl.lo1 ^ xl) & (pl.hi1 ^ xh) & 0x1FFF1FFFu;
}

// Test of the constrained seats: their ranges (what & TEST_RANGES), their rule programs (what & TEST_PROGS).
// `what` carries the seats to look at in bits 4..7 (TEST_SEATS(mask)); TEST_ALL = everything of every seat.
constexpr uint32_t TEST_RANGES = 1u, TEST_PROGS = 2u, TEST_ALL = 0xF3u;
__host__ __device__ constexpr uint32_t TEST_SEATS(uint32_t mask) { return (mask & 0xFu) << 4; }
__device__ __forceinline__ bool seats_pass(const SampleArgs &a, const Planes &pl, uint32_t what, uint32_t *scratch, unsigned lane)
{
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const SeatC &c = a.cons.seat[k];
        if (c.active && ((what >> (4 + k)) & 1u) && ((what & TEST_RANGES) || c.prog != 0xFFFFFFFFu)) {
            uint32_t m0, m1;
            seat_words(pl, k, m0, m1);
            const Feat f = hand_features(m0, m1);
            if (what & TEST_RANGES) ok = ok && seat_ranges_ok(c, f);
            if ((what & TEST_PROGS) && c.prog != 0xFFFFFFFFu) {
                bool err = false;
#ifdef BMC_JIT
                const bool v = bmc_jit_seat_rule(k, f, err);        // the seat's rule form, compiled to native code
#else
                const bool v = run_program(a.cons.code + c.prog, feat_store(scratch, lane, f), err);
#endif
                ok = ok && v && !err;
            }
        }
    }
    return ok;
}

// Stage B: test of the constrained seats (`what`: ranges, programs or both), then tallies and the compacted store.
__device__ __forceinline__ void stage_b(const SampleArgs &a, const Planes &pl, bool active, uint32_t *hist,
                                        const uint8_t *shape_lut, unsigned lane, uint32_t *scratch, uint32_t what = TEST_ALL)
{
    const bool ok = seats_pass(a, pl, what, scratch, lane) && active;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
    if (bal == 0u) return;
    const unsigned cnt = __popc(bal);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&a.tallies[2], (unsigned long long)cnt);      // accepted (also the slot cursor)
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (ok) {
        const unsigned long long slot = base + __popc(bal & ((1u << lane) - 1u));
        if (a.records != nullptr && slot < a.cap) a.records[slot] = make_uint4(pl.lo0, pl.lo1, pl.hi0, pl.hi1);
        if (a.tally_mask) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t m0, m1;
                seat_words(pl, k, m0, m1);
                const Feat f = hand_features(m0, m1);
                if (a.tally_mask & BMC_TALLY_HCP) atomicAdd(&hist[SH_HCP + k * 40 + (f.f2 & 0xFFu)], 1u);
                if (a.tally_mask & BMC_TALLY_LEN) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) atomicAdd(&hist[SH_LEN + (k * 4 + s) * 16 + ((f.f0 >> (8 * s)) & 0xFFu)], 1u);
                }
                if (a.tally_mask & BMC_TALLY_SHAPE) {
                    const uint32_t L = f.f0;
                    const uint32_t cls = shape_lut[(L & 0xFFu) * 196u + ((L >> 8) & 0xFFu) * 14u + ((L >> 16) & 0xFFu)];
                    atomicAdd(&hist[SH_SHAPE + k * 40 + cls], 1u);
                }
            }
        }
    }
}

// block-level flush of the per-warp shared histograms into bmc_tallies (u64 atomics)
__device__ __forceinline__ void flush_hist(const SampleArgs &a, const uint32_t *s_hist, int copies = WARPS_PER_BLOCK)
{
    if (!a.tally_mask) return;
    for (int i = threadIdx.x; i < SH_WORDS; i += THREADS) {
        uint32_t v = 0;
#pragma unroll
        for (int w = 0; w < copies; ++w) v += s_hist[w * SH_WORDS + i];
        if (v == 0u) continue;
        int word;                                                                    // padded shared layout -> bmc_tallies words
        if (i < SH_LEN) word = 4 + i;                                                // hcp[4][40]
        else if (i < SH_SHAPE) {
            const int j = i - SH_LEN, bin = j & 15;
            if (bin >= BMC_LEN_BINS) continue;
            word = 4 + 160 + (j >> 4) * BMC_LEN_BINS + bin;                          // len[4][4][14]
        } else word = 4 + 160 + 224 + (i - SH_SHAPE);                                // shape[4][40]
        atomicAdd(&a.tallies[word], (unsigned long long)v);
    }
}

template <bool HAS_CONS, bool FIXED>
__device__ __forceinline__ void sample_body(const SampleArgs &a)
{
    __shared__ uint32_t s_hist[WARPS_PER_BLOCK][SH_WORDS];
    __shared__ uint4 s_queue[HAS_CONS ? WARPS_PER_BLOCK : 1][QCAP];
    __shared__ uint32_t s_feat[WARPS_PER_BLOCK][FEAT_SCRATCH_WORDS];
    __shared__ uint8_t s_shape[14 * 14 * 14];

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < WARPS_PER_BLOCK * SH_WORDS; i += THREADS) (&s_hist[0][0])[i] = 0u;
    for (int i = threadIdx.x; i < 14 * 14 * 14; i += THREADS) s_shape[i] = c_shape_lut[i];
    __syncthreads();

    uint32_t *hist = s_hist[warp];
    const unsigned long long total_warps = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    uint32_t n_raw = 0, n_void = 0;           // per thread: at most `iters` each
    unsigned q_head = 0, q_tail = 0;          // warp-uniform

    for (uint32_t it = 0; it < a.iters; ++it) {
        const unsigned long long off = ((unsigned long long)it * total_warps + gwarp) * 32ull + lane;
        const bool valid = off < a.n;
        const unsigned long long idx = a.first + off;
        Planes pl;
        const bool voided = FIXED ? deal_fixed(a.key, a.fix, (uint32_t)idx, (uint32_t)(idx >> 32), pl)
                                  : deal52(a.key, (uint32_t)idx, (uint32_t)(idx >> 32), pl);
        n_raw += valid ? 1u : 0u;
        n_void += (valid && voided) ? 1u : 0u;
        bool ok = valid && !voided;
        if (HAS_CONS) {
            // Stage A: cheap range test of the primary seat
            const SeatC &c = a.cons.seat[a.cons.primary];
            uint32_t m0, m1;
            seat_words(pl, (int)a.cons.primary, m0, m1);
            const Feat f = hand_features(m0, m1);
            ok = ok && seat_ranges_ok(c, f);
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
            if (bal) {
                if (ok) {
                    const unsigned pos = (q_tail + __popc(bal & ((1u << lane) - 1u))) % QCAP;
                    s_queue[warp][pos] = make_uint4(pl.lo0, pl.lo1, pl.hi0, pl.hi1);
                }
                q_tail += __popc(bal);
                __syncwarp();
                if (q_tail - q_head >= 32u) {
                    const uint4 r = s_queue[warp][(q_head + lane) % QCAP];
                    q_head += 32u;
                    __syncwarp();
                    Planes p2 = {r.x, r.y, r.z, r.w};
                    stage_b(a, p2, true, hist, s_shape, lane, s_feat[warp]);
                }
            }
        } else {
            stage_b(a, pl, ok, hist, s_shape, lane, s_feat[warp]);
        }
    }
    if (HAS_CONS) {
        const unsigned left = q_tail - q_head;       // < 32
        if (left) {
            const uint4 r = s_queue[warp][(q_head + lane) % QCAP];
            Planes p2 = {r.x, r.y, r.z, r.w};
            stage_b(a, p2, lane < left, hist, s_shape, lane, s_feat[warp]);
        }
    }

    // flush counters and histograms
    {
        unsigned long long r64 = n_raw, v64 = n_void;
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
    flush_hist(a, &s_hist[0][0]);
}

// ---------------------------------------------------------------------------
// K1b: seat-first sampler (seatfirst.cuh): stages A1 -> A2 -> B1 -> B2 with per-warp
// compaction queues between them, so each stage runs on full warps.
// ---------------------------------------------------------------------------
#ifndef SF_MIN_BLOCKS
#define SF_MIN_BLOCKS 6
#endif
constexpr int SF_HIST_COPIES = 2;
#ifndef SF_CALLS
#define SF_CALLS 4            /* Philox calls (four indices each) per lane and iteration of stage A1 */
#endif
constexpr int SF_QCAP1 = 128; // stage A1 -> A2 queue (32-bit entries); rounds that do not fit wait for a drain
__constant__ LowTabs c_lowtabs = make_lowtabs();
__constant__ uint32_t c_thresh39[10] = {thresh39(0), thresh39(1), thresh39(2), thresh39(3), thresh39(4),
                                        thresh39(5), thresh39(6), thresh39(7), thresh39(8), thresh39(9)};

__device__ __forceinline__ void sample_seatfirst_body(const SampleArgs &a)
{
    __shared__ uint32_t s_hist[SF_HIST_COPIES][SH_WORDS];   // stage B is rare here: few copies suffice
    __shared__ uint32_t s_q1[WARPS_PER_BLOCK][SF_QCAP1];   // A1 survivors: index - a.first
    __shared__ uint4 s_q2[WARPS_PER_BLOCK][QCAP];     // A2 survivors: (m0, m1, index - a.first, -)
    __shared__ uint4 s_q3[WARPS_PER_BLOCK][QCAP];     // B1 survivors: complete deals (planes)
    __shared__ uint32_t s_feat[WARPS_PER_BLOCK][FEAT_SCRATCH_WORDS];
    __shared__ uint8_t s_shape[14 * 14 * 14];
    __shared__ __align__(16) LowTabs s_low;

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SF_HIST_COPIES * SH_WORDS; i += THREADS) (&s_hist[0][0])[i] = 0u;
    for (int i = threadIdx.x; i < 14 * 14 * 14; i += THREADS) s_shape[i] = c_shape_lut[i];
    for (int i = threadIdx.x; i < (int)(sizeof(LowTabs) / 4); i += THREADS)
        reinterpret_cast<uint32_t *>(&s_low)[i] = reinterpret_cast<const uint32_t *>(&c_lowtabs)[i];
    __syncthreads();

    uint32_t *hist = s_hist[warp % SF_HIST_COPIES];
    uint4 *q2 = s_q2[warp], *q3 = s_q3[warp];
    unsigned h3 = 0, t3 = 0;
    // Stage B in two steps with a compaction between them, when there is something to gain:
    //   interpreted rule programs: B1 = the cheap range tests of every seat, B2 = the programs, on full warps of
    //                              deals that can still be accepted;
    //   compiled rule forms (NVRTC build): B1 = ranges and rules of the primary and the most selective other
    //                              seat (b1_mask, calibrated), B2 = the remaining seats.
#ifdef BMC_JIT
    const uint32_t rest = ~a.cons.b1_mask & 0xFu;
    const bool two_stage = rest != 0u;
    const uint32_t what1 = TEST_SEATS(a.cons.b1_mask) | TEST_RANGES | TEST_PROGS, what2 = TEST_SEATS(rest) | TEST_RANGES | TEST_PROGS;
#else
    const bool two_stage = a.cons.n_prog != 0u;
    const uint32_t what1 = TEST_SEATS(0xFu) | TEST_RANGES, what2 = TEST_SEATS(0xFu) | TEST_PROGS;
#endif
    const unsigned long long total_warps = (unsigned long long)gridDim.x * WARPS_PER_BLOCK;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    uint32_t n_void = 0;                             // per thread
    unsigned h1 = 0, t1 = 0, h2 = 0, t2 = 0;         // queue cursors, warp-uniform
    const uint32_t P = a.cons.primary;
    const SeatC &c = a.cons.seat[P];

    // ---- stage B on up to 32 queued hands ------------------------------------------
    auto run_b = [&](unsigned count) {
        const uint4 r = q2[(h2 + lane) % QCAP];       // (m0, m1, index - a.first, -)
        h2 += count;
        __syncwarp();
        const bool active = lane < count;
        Planes pl;
        const uint32_t xl = (P & 1u) ? 0xFFFFFFFFu : 0u, xh = (P & 2u) ? 0xFFFFFFFFu : 0u;
        const unsigned long long idx = a.first + r.z;
        const bool voided = deal_rest39(a.key, a.sf_q0, r.x, r.y, xl, xh, (uint32_t)idx, (uint32_t)(idx >> 32), pl);
        n_void += (active && voided) ? 1u : 0u;
        if (!two_stage) {
            stage_b(a, pl, active && !voided, hist, s_shape, lane, s_feat[warp], TEST_ALL);
            return;
        }
        const bool ok = seats_pass(a, pl, what1, s_feat[warp], lane) && active && !voided;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (bal) {
            if (ok) q3[(t3 + __popc(bal & ((1u << lane) - 1u))) % QCAP] = make_uint4(pl.lo0, pl.lo1, pl.hi0, pl.hi1);
            t3 += __popc(bal);
            __syncwarp();
            if (t3 - h3 >= 32u) {
                const uint4 r3 = q3[(h3 + lane) % QCAP];
                h3 += 32u;
                __syncwarp();
                const Planes p3 = {r3.x, r3.y, r3.z, r3.w};
                stage_b(a, p3, true, hist, s_shape, lane, s_feat[warp], what2);
            }
        }
    };
    // ---- stage A2 on up to 32 queued survivors of the HCP test ----------------------------
    auto run_a2 = [&](unsigned count) {
        const uint32_t off = s_q1[warp][(h1 + lane) % SF_QCAP1];      // index - a.first
        h1 += count;
        __syncwarp();
        const bool active = lane < count;
        const unsigned long long idx = a.first + off;
        uint32_t wh[4], wl[4];
        philox4x32_10(a.key, (uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, STREAM_SEATFIRST_RANK, wh);
        philox4x32_10(a.key, (uint32_t)(idx >> 2), (uint32_t)(idx >> 34), 0u, STREAM_SEATFIRST_RANK_LO, wl);
        const uint32_t j = (uint32_t)idx & 3u;
        const uint32_t u_hi = j < 2u ? (j == 0u ? wh[0] : wh[1]) : (j == 2u ? wh[2] : wh[3]);
        const uint32_t u_lo = j < 2u ? (j == 0u ? wl[0] : wl[1]) : (j == 2u ? wl[2] : wl[3]);
        unsigned long long d;
        uint32_t magic, i, ell, m0, m1;
        // lanes past the end of the queue look up u = 0: class 0, nothing to scan
        const uint32_t info = sf_find_class(a.cons.sf_tab, a.cons.sf_bmul, active ? u_lo : 0u, active ? u_hi : 0u, d, magic);
        sf_split(d, (info >> 12) & 0x7FFu, magic, i, ell);
        const uint32_t H = sf_honours_of(info, i);
        sf_unrank_lows(s_low, ell, 13u - (uint32_t)__popc(H), m0, m1);
        m0 |= ((H & 0xFu) << 9) | ((H & 0xF0u) << 21);
        m1 |= ((H & 0xF00u) << 1) | ((H & 0xF000u) << 13);
        const Feat f = hand_features(m0, m1);
        const bool ok = active && seat_ranges_ok(c, f);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (bal) {
            if (ok) q2[(t2 + __popc(bal & ((1u << lane) - 1u))) % QCAP] = make_uint4(m0, m1, off, 0u);
            t2 += __popc(bal);
            __syncwarp();
        }
    };

    // Stage A1.  The uniform of index i is u = (hi word << 32) | lo word, the words (i & 3) of the Philox
    // calls with counter i >> 2 on streams 4 (hi) and 5 (lo).  u < theta is decided by the hi word alone
    // unless it EQUALS theta's hi word (2^-32 per index), so one Philox call tests four indices; the lo
    // call is made for ties, for the void check (u >= S C(52,13), hi word >= 0xFFFFFFCB) and by stage A2.
    // Per iteration a lane tests SF_CALLS calls = 4 SF_CALLS indices and keeps one survivor bit per index;
    // the survivors are queued as 32-bit offsets from a.first (a launch covers at most 2^31 indices,
    // host), one per lane and round.  Rounds that do not fit wait for a drain.
    const uint32_t tid_g = (uint32_t)gwarp * 32u + lane;
    const uint32_t stride = (uint32_t)total_warps * 32u;                       // calls per pass of the grid
    const unsigned long long base_call = a.first >> 2;
    const uint32_t mis = (uint32_t)a.first & 3u, n32 = (uint32_t)a.n;
    const uint32_t full_iters = n32 / (4u * SF_CALLS * stride);                // iterations before it: every index valid
    const uint32_t theta_hi = (uint32_t)(a.cons.sf_theta >> 32), void_hi = (uint32_t)(a.cons.sf_void >> 32);
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t it = 0, mask = 0, co0 = 0;
    bool pending = false;
    for (;;) {
        if (!pending && it < a.iters) {
            co0 = it * (SF_CALLS * stride) + tid_g;                            // call offset of bits 0..3; bits 4k..4k+3: + k * stride
            mask = 0;
            uint32_t top = 0u;
            bool tie = false;
#pragma unroll
            for (int k = 0; k < SF_CALLS; ++k) {
                uint32_t w[4];
                const unsigned long long call = base_call + (co0 + (uint32_t)k * stride);
                philox4x32_10(a.key, (uint32_t)call, (uint32_t)(call >> 32), 0u, STREAM_SEATFIRST_RANK, w);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (w[j] < theta_hi) mask |= 1u << (4 * k + j);
                    tie = tie || w[j] == theta_hi;
                    top = max(top, w[j]);
                }
            }
            if (__any_sync(0xFFFFFFFFu, tie || top >= void_hi) || it == 0u || it >= full_iters) {
                // edge iterations: drop the indices outside [first, first + n); rare: ties and the void check
#pragma unroll 1
                for (int k = 0; k < SF_CALLS; ++k) {
                    uint32_t wh[4], wl[4];
                    const uint32_t co = co0 + (uint32_t)k * stride;
                    const unsigned long long call = base_call + co;
                    philox4x32_10(a.key, (uint32_t)call, (uint32_t)(call >> 32), 0u, STREAM_SEATFIRST_RANK, wh);
                    philox4x32_10(a.key, (uint32_t)call, (uint32_t)(call >> 32), 0u, STREAM_SEATFIRST_RANK_LO, wl);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool valid = 4u * co + (uint32_t)j - mis < n32;      // wraps for the indices before a.first
                        const unsigned long long u = (((unsigned long long)wh[j]) << 32) | wl[j];
                        const uint32_t bit = 1u << (4 * k + j);
                        mask = (valid && u < a.cons.sf_theta) ? (mask | bit) : (mask & ~bit);
                        if (valid && u >= a.cons.sf_void) ++n_void;
                    }
                }
            }
            ++it;
        }
        for (;;) {                                                             // queue one survivor per lane and round
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, mask != 0u);
            pending = false;
            if (bal == 0u) break;
            pending = true;
            if (t1 - h1 + __popc(bal) > (unsigned)SF_QCAP1) break;
            if (mask != 0u) {
                const uint32_t b = (uint32_t)__ffs((int)mask) - 1u;
                mask &= mask - 1u;
                s_q1[warp][(t1 + __popc(bal & lt)) % SF_QCAP1] = 4u * (co0 + (b >> 2) * stride) + (b & 3u) - mis;
            }
            t1 += __popc(bal);
        }
        __syncwarp();
        const bool fin = !pending && it >= a.iters;
        // drain the queues: full warps while the stream lasts, whatever is left at the end
        for (;;) {
            const unsigned n1 = t1 - h1;
            const bool do_a2 = n1 >= 32u || (fin && n1 != 0u);
            if (do_a2) run_a2(n1 < 32u ? n1 : 32u);
            const unsigned n2 = t2 - h2;
            const bool do_b = n2 >= 32u || (fin && n2 != 0u && t1 == h1);
            if (do_b) run_b(n2 < 32u ? n2 : 32u);
            if (!do_a2 && !do_b) break;
        }
        if (fin) break;
    }
    if (t3 - h3) {
        const uint4 r3 = q3[(h3 + lane) % QCAP];
        const Planes p3 = {r3.x, r3.y, r3.z, r3.w};
        stage_b(a, p3, lane < t3 - h3, hist, s_shape, lane, s_feat[warp], what2);
    }

    {
        unsigned long long v64 = n_void;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v64 += __shfl_xor_sync(0xFFFFFFFFu, v64, o);
        if (lane == 0 && v64) atomicAdd(&a.tallies[1], v64);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.tallies[0], a.n);   // the grid covers exactly [first, first + n)
    }
    __syncthreads();
    flush_hist(a, &s_hist[0][0], SF_HIST_COPIES);
}


#ifndef BMC_JIT
template <bool HAS_CONS, bool FIXED>
__global__ void __launch_bounds__(THREADS) sample_kernel(const __grid_constant__ SampleArgs a)
{
    sample_body<HAS_CONS, FIXED>(a);
}
__global__ void __launch_bounds__(THREADS, SF_MIN_BLOCKS) sample_seatfirst_kernel(const __grid_constant__ SampleArgs a)
{
    sample_seatfirst_body(a);
}
#else
#ifdef BMC_JIT_FULL
extern "C" __global__ void __launch_bounds__(THREADS) bmc_jit_sample_full(const __grid_constant__ SampleArgs a)
{
    sample_body<true, false>(a);
}
#endif
#ifdef BMC_JIT_SEATFIRST
extern "C" __global__ void __launch_bounds__(THREADS, SF_MIN_BLOCKS) bmc_jit_sample_seatfirst(const __grid_constant__ SampleArgs a)
{
    sample_seatfirst_body(a);
}
#endif
#endif
