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
#include "score.cuh"

#ifndef BMC_TALLY_HCP          /* NVRTC build: the few constants of bridgemc.h the kernels need */
#define BMC_TALLY_HCP   1u
#define BMC_TALLY_LEN   2u
#define BMC_TALLY_SHAPE 4u
#define BMC_TALLY_SCORE 8u
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
    unsigned long long sf_theta;   // u < sf_theta: the honour holding has an admissible HCP (mode 4: the hand satisfies the seat)
    unsigned long long sf_void;    // u >= sf_void = S * C(52,13) voids the index (probability 1.2e-8)
    ExactTab x;                    // mode 4: exact class table of the primary seat (distgen.cuh)
};

// shared-memory histogram (per copy): hcp[4][40] | len[4][4][16] | shape[4][40]
constexpr int SH_HCP = 0, SH_LEN = 160, SH_SHAPE = 160 + 256, SH_WORDS = 160 + 256 + 160;
constexpr int WARPS_PER_BLOCK = 8, THREADS = WARPS_PER_BLOCK * 32;
constexpr int QCAP = 64;     // per-warp queue of stage-A survivors
constexpr int FULL_HIST_COPIES = 8, SF_HIST_COPIES = 2;
// words of bmc_tallies: 4 counters | hcp[4][40] | len[4][4][14] | shape[4][40] | tricks[8][14] | imps[8][49] | imp_dropped[8]
constexpr int TALLY_SCORE_WORD = 4 + 160 + 224 + 160;

// (len_c, len_d, len_h) -> sorted-shape class: the 39 shapes a >= b >= c >= d, a + b + c + d = 13, in the order
// of bmc_shape_table (a ascending from 4, then b, then c).  A compile-time constant (no upload).
struct ShapeLut { uint8_t v[14 * 14 * 14]; };
struct ShapeList { uint8_t s[39][4]; };
__host__ __device__ constexpr ShapeList make_shape_list()
{
    ShapeList L{};
    int n = 0;
    for (int a = 4; a <= 13; ++a)
        for (int b = 0; b <= a; ++b)
            for (int c = 0; c <= b; ++c) {
                const int d = 13 - a - b - c;
                if (d < 0 || d > c) continue;
                L.s[n][0] = (uint8_t)a; L.s[n][1] = (uint8_t)b; L.s[n][2] = (uint8_t)c; L.s[n][3] = (uint8_t)d;
                ++n;
            }
    return L;
}
__host__ __device__ constexpr ShapeLut make_shape_lut()
{
    ShapeLut T{};
    const ShapeList L = make_shape_list();
    for (int i = 0; i < 14 * 14 * 14; ++i) T.v[i] = 0xFF;
    for (int x = 0; x <= 13; ++x)
        for (int y = 0; x + y <= 13; ++y)
            for (int z = 0; x + y + z <= 13; ++z) {
                int v[4] = {x, y, z, 13 - x - y - z};
                for (int i = 0; i < 4; ++i)
                    for (int j = i + 1; j < 4; ++j)
                        if (v[j] > v[i]) { const int t = v[i]; v[i] = v[j]; v[j] = t; }
                for (int k = 0; k < 39; ++k)
                    if (L.s[k][0] == v[0] && L.s[k][1] == v[1] && L.s[k][2] == v[2] && L.s[k][3] == v[3])
                        T.v[x * 196 + y * 14 + z] = (uint8_t)k;
            }
    return T;
}
__constant__ ShapeLut c_shape_lut = make_shape_lut();
constexpr int SHAPE_LUT_WORDS = (14 * 14 * 14 + 3) / 4;

struct SampleArgs {
    ConsDev cons;
    PhiloxKey key;
    FixDev fix;
    uint64_t first, n;
    void *records;           // may be null; uint4 records, or 32-bit index offsets when rec_mode = 1
    uint64_t cap;
    unsigned long long *tallies;   // bmc_tallies as u64 words
    uint32_t tally_mask;
    uint32_t iters;          // warp-iterations per warp (seat-first: each covers 4 SF_CALLS indices per lane)
    uint32_t sf_q0;          // seat-first mode: initial Q of stage B (capacity 0 for the primary seat)
    uint32_t rec_mode;       // 0: 128-bit records; 1: the accepted deal's index as a 32-bit offset from the CALL's first index
    uint32_t idx_base;       // rec_mode 1: a.first - (first index of the call)
    uint32_t flags;          // bit 0: this launch interprets rule programs (needs the feature scratch)
    uint64_t target;         // != 0: the launch does nothing if *snap - start_acc >= target: accept-target jobs enqueue their
    uint64_t start_acc;      //       waves back to back and the waves after the one that reaches the target are no-ops.
    const unsigned long long *snap;   // the accepted counter as it stood when the previous wave ended (publish_counters_kernel)
    ContractTable score;     // BMC_TALLY_SCORE: contracts / IMP pairs tallied over the accepted deals (score.cuh)
    PhiloxKey tkey;          // key of the stand-in trick source
};
constexpr uint32_t ARGS_INTERP = 1u;

// Dynamic shared memory: only what the launch needs, so the common configurations keep their occupancy.
// Offsets in 32-bit words from the start of the dynamic segment; computed identically by host and device.
struct SmemPlan {
    uint32_t low, q1, q2, q3, q3i, shape, hist, feat, sc_hist, sc_tab, words;
};
__host__ __device__ inline uint32_t smem_take(uint32_t &cur, uint32_t words)
{
    const uint32_t at = cur;
    cur += (words + 3u) & ~3u;                  // 16-byte granules
    return at;
}
// seatfirst = the seat-first kernel; two_stage = stage B is split by a third queue
__host__ __device__ inline SmemPlan smem_plan(bool seatfirst, bool has_queue, bool two_stage, bool interp, uint32_t tally_mask)
{
    SmemPlan p{};
    uint32_t cur = 0;
    p.low = seatfirst ? smem_take(cur, (uint32_t)(sizeof(LowTabs) / 4)) : 0u;
    p.q2 = has_queue ? smem_take(cur, WARPS_PER_BLOCK * QCAP * 4) : 0u;                 // uint4 entries
    p.q1 = has_queue ? smem_take(cur, WARPS_PER_BLOCK * (seatfirst ? 128 : QCAP)) : 0u;  // seat-first: A1 survivors; full deal: offsets of q2's entries
    p.q3 = two_stage ? smem_take(cur, WARPS_PER_BLOCK * QCAP * 4) : 0u;
    p.q3i = two_stage ? smem_take(cur, WARPS_PER_BLOCK * QCAP) : 0u;
    p.shape = (tally_mask & BMC_TALLY_SHAPE) ? smem_take(cur, SHAPE_LUT_WORDS) : 0u;
    p.hist = (tally_mask & 7u) ? smem_take(cur, (seatfirst ? SF_HIST_COPIES : FULL_HIST_COPIES) * SH_WORDS) : 0u;
    p.feat = interp ? smem_take(cur, WARPS_PER_BLOCK * FEAT_SCRATCH_WORDS) : 0u;
    p.sc_hist = (tally_mask & BMC_TALLY_SCORE) ? smem_take(cur, SCORE_WORDS) : 0u;
    p.sc_tab = (tally_mask & BMC_TALLY_SCORE) ? smem_take(cur, SCORE_TR_WORDS) : 0u;
    p.words = cur;
    return p;
}

// what stage B needs of the block's shared memory
struct TallyMem {
    uint32_t *hist;          // this warp's histogram copy
    const uint8_t *shape;    // sorted-shape LUT
    uint32_t *sc_hist;       // score histogram (one per block)
    const int *sc_tab;       // base-score table [8][14]
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
// `off` = the deal's index - a.first (stored instead of the record when rec_mode = 1).
__device__ __forceinline__ void stage_b(const SampleArgs &a, const Planes &pl, uint32_t off, bool active, const TallyMem &tm,
                                        unsigned lane, uint32_t *scratch, uint32_t what = TEST_ALL)
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
        if (a.records != nullptr && slot < a.cap) {
            if (a.rec_mode == 0u) reinterpret_cast<uint4 *>(a.records)[slot] = make_uint4(pl.lo0, pl.lo1, pl.hi0, pl.hi1);
            else reinterpret_cast<uint32_t *>(a.records)[slot] = a.idx_base + off;
        }
        if (a.tally_mask & 7u) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t m0, m1;
                seat_words(pl, k, m0, m1);
                const Feat f = hand_features(m0, m1);
                if (a.tally_mask & BMC_TALLY_HCP) atomicAdd(&tm.hist[SH_HCP + k * 40 + (f.f2 & 0xFFu)], 1u);
                if (a.tally_mask & BMC_TALLY_LEN) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) atomicAdd(&tm.hist[SH_LEN + (k * 4 + s) * 16 + ((f.f0 >> (8 * s)) & 0xFFu)], 1u);
                }
                if (a.tally_mask & BMC_TALLY_SHAPE) {
                    const uint32_t L = f.f0;
                    const uint32_t cls = tm.shape[(L & 0xFFu) * 196u + ((L >> 8) & 0xFFu) * 14u + ((L >> 16) & 0xFFu)];
                    atomicAdd(&tm.hist[SH_SHAPE + k * 40 + cls], 1u);
                }
            }
        }
        if (a.tally_mask & BMC_TALLY_SCORE) {
            // compscore tallies over the accepted deal (base-score / match-points, compscore.cl:94-131): the stand-in
            // trick source is keyed on the deal itself (score.cuh)
            uint32_t tr[8];
            tricks_of(a.tkey, pl.lo0, pl.lo1, pl.hi0, pl.hi1, tr);
            score_tally_update(a.score, tm.sc_tab, tr, tm.sc_hist);
        }
    }
}

// block-level set-up / flush of the shared histograms
__device__ __forceinline__ void tally_init(const SampleArgs &a, const SmemPlan &p, uint32_t *smem, int copies)
{
    if (a.tally_mask & 7u)
        for (int i = threadIdx.x; i < copies * SH_WORDS; i += THREADS) smem[p.hist + i] = 0u;
    if (a.tally_mask & BMC_TALLY_SHAPE)
        for (int i = threadIdx.x; i < SHAPE_LUT_WORDS; i += THREADS)
            smem[p.shape + i] = reinterpret_cast<const uint32_t *>(c_shape_lut.v)[i];
    if (a.tally_mask & BMC_TALLY_SCORE) {
        for (int i = threadIdx.x; i < SCORE_WORDS; i += THREADS) smem[p.sc_hist + i] = 0u;
        for (int i = threadIdx.x; i < SCORE_TR_WORDS; i += THREADS) {
            const int c = i / 14;
            reinterpret_cast<int *>(smem)[p.sc_tab + i] = c < (int)a.score.n ? base_score_dev(a.score.c[c], i % 14, a.score.vul[c]) : 0;
        }
    }
}
__device__ __forceinline__ TallyMem tally_mem(const SmemPlan &p, uint32_t *smem, unsigned copy)
{
    TallyMem tm;
    tm.hist = smem + p.hist + copy * SH_WORDS;
    tm.shape = reinterpret_cast<const uint8_t *>(smem + p.shape);
    tm.sc_hist = smem + p.sc_hist;
    tm.sc_tab = reinterpret_cast<const int *>(smem + p.sc_tab);
    return tm;
}
__device__ __forceinline__ void flush_hist(const SampleArgs &a, const SmemPlan &p, const uint32_t *smem, int copies)
{
    if (a.tally_mask & 7u) {
        const uint32_t *s_hist = smem + p.hist;
        for (int i = threadIdx.x; i < SH_WORDS; i += THREADS) {
            uint32_t v = 0;
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
    if (a.tally_mask & BMC_TALLY_SCORE)
        for (int i = threadIdx.x; i < SCORE_WORDS; i += THREADS) {
            const uint32_t v = smem[p.sc_hist + i];
            if (v) atomicAdd(&a.tallies[TALLY_SCORE_WORD + i], (unsigned long long)v);
        }
}

// accept-target jobs: a wave launched after the target was reached does nothing (and counts nothing).  The test
// reads a snapshot taken between the waves -- not the live counter, which this very wave advances -- so every
// block of a wave takes the same decision and the accepted set stays "whole waves until the target is met".
__device__ __forceinline__ bool wave_cancelled(const SampleArgs &a)
{
    return a.target != 0ull && __ldg(a.snap) - a.start_acc >= a.target;
}

extern __shared__ __align__(16) uint32_t s_dyn[];

template <bool HAS_CONS, bool FIXED>
__device__ __forceinline__ void sample_body(const SampleArgs &a)
{
    if (wave_cancelled(a)) return;
    const SmemPlan sp = smem_plan(false, HAS_CONS, false, (a.flags & ARGS_INTERP) != 0u, a.tally_mask);
    uint32_t *smem = s_dyn;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    tally_init(a, sp, smem, FULL_HIST_COPIES);
    __syncthreads();

    const TallyMem tm = tally_mem(sp, smem, warp);
    uint4 *queue = reinterpret_cast<uint4 *>(smem + sp.q2) + warp * QCAP;
    uint32_t *qoff = smem + sp.q1 + warp * QCAP;
    uint32_t *feat = smem + sp.feat + warp * FEAT_SCRATCH_WORDS;
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
                    queue[pos] = make_uint4(pl.lo0, pl.lo1, pl.hi0, pl.hi1);
                    qoff[pos] = (uint32_t)off;
                }
                q_tail += __popc(bal);
                __syncwarp();
                if (q_tail - q_head >= 32u) {
                    const uint4 r = queue[(q_head + lane) % QCAP];
                    const uint32_t o = qoff[(q_head + lane) % QCAP];
                    q_head += 32u;
                    __syncwarp();
                    Planes p2 = {r.x, r.y, r.z, r.w};
                    stage_b(a, p2, o, true, tm, lane, feat);
                }
            }
        } else {
            stage_b(a, pl, (uint32_t)off, ok, tm, lane, feat);
        }
    }
    if (HAS_CONS) {
        const unsigned left = q_tail - q_head;       // < 32
        if (left) {
            const uint4 r = queue[(q_head + lane) % QCAP];
            const uint32_t o = qoff[(q_head + lane) % QCAP];
            Planes p2 = {r.x, r.y, r.z, r.w};
            stage_b(a, p2, o, lane < left, tm, lane, feat);
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
    flush_hist(a, sp, smem, FULL_HIST_COPIES);
}

// ---------------------------------------------------------------------------
// K1b: seat-first sampler (seatfirst.cuh): stages A1 -> A2 -> B1 -> B2 with per-warp
// compaction queues between them, so each stage runs on full warps.
// ---------------------------------------------------------------------------
#ifndef SF_MIN_BLOCKS
#define SF_MIN_BLOCKS 6
#endif
#ifndef SF_CALLS
#define SF_CALLS 4            /* Philox calls (four indices each) per lane and iteration of stage A1 */
#endif
constexpr int SF_QCAP1 = 128; // stage A1 -> A2 queue (32-bit entries); rounds that do not fit wait for a drain
__constant__ LowTabs c_lowtabs = make_lowtabs();

// Stage B in two steps with a compaction between them, when there is something to gain:
//   interpreted rule programs: B1 = the cheap range tests of every seat, B2 = the programs, on full warps of
//                              deals that can still be accepted;
//   compiled rule forms (NVRTC build): B1 = ranges and rules of the primary and the most selective other
//                              seat (b1_mask, calibrated), B2 = the remaining seats.
__host__ __device__ inline bool sf_two_stage(const ConsDev &c, bool jit)
{
    return jit ? (~c.b1_mask & 0xFu) != 0u : c.n_prog != 0u;
}

// EXACT = mode 4: stage A1 accepts exactly the primary seat's admissible hands (class table), stage A2 unranks
// without testing and stage B leaves the primary seat alone.
template <bool EXACT>
__device__ __forceinline__ void sample_seatfirst_body(const SampleArgs &a)
{
    if (wave_cancelled(a)) return;
#ifdef BMC_JIT
    constexpr bool kJit = true;
#else
    constexpr bool kJit = false;
#endif
    const bool two_stage = sf_two_stage(a.cons, kJit);
    const SmemPlan sp = smem_plan(true, true, two_stage, (a.flags & ARGS_INTERP) != 0u, a.tally_mask);
    uint32_t *smem = s_dyn;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    tally_init(a, sp, smem, SF_HIST_COPIES);
    for (int i = threadIdx.x; i < (int)(sizeof(LowTabs) / 4); i += THREADS)
        smem[sp.low + i] = reinterpret_cast<const uint32_t *>(&c_lowtabs)[i];
    __syncthreads();

    const LowTabs &s_low = *reinterpret_cast<const LowTabs *>(smem + sp.low);
    const TallyMem tm = tally_mem(sp, smem, warp % SF_HIST_COPIES);
    uint32_t *q1 = smem + sp.q1 + warp * SF_QCAP1;                                  // A1 survivors: index - a.first
    uint4 *q2 = reinterpret_cast<uint4 *>(smem + sp.q2) + warp * QCAP;              // A2 survivors: (m0, m1, index - a.first, -)
    uint4 *q3 = reinterpret_cast<uint4 *>(smem + sp.q3) + warp * QCAP;              // B1 survivors: complete deals (planes)
    uint32_t *q3i = smem + sp.q3i + warp * QCAP;                                    //               and their index - a.first
    uint32_t *feat = smem + sp.feat + warp * FEAT_SCRATCH_WORDS;
    unsigned h3 = 0, t3 = 0;
    const uint32_t seats = EXACT ? 0xFu & ~(1u << a.cons.primary) : 0xFu;      // mode 4: the primary seat is admitted by construction
#ifdef BMC_JIT
    const uint32_t rest = ~a.cons.b1_mask & 0xFu;
    const uint32_t what1 = TEST_SEATS(a.cons.b1_mask & seats) | TEST_RANGES | TEST_PROGS, what2 = TEST_SEATS(rest & seats) | TEST_RANGES | TEST_PROGS;
#else
    const uint32_t what1 = TEST_SEATS(seats) | TEST_RANGES, what2 = TEST_SEATS(seats) | TEST_PROGS;
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
            stage_b(a, pl, r.z, active && !voided, tm, lane, feat, TEST_SEATS(seats) | TEST_RANGES | TEST_PROGS);
            return;
        }
        const bool ok = seats_pass(a, pl, what1, feat, lane) && active && !voided;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (bal) {
            if (ok) {
                const unsigned pos = (t3 + __popc(bal & ((1u << lane) - 1u))) % QCAP;
                q3[pos] = make_uint4(pl.lo0, pl.lo1, pl.hi0, pl.hi1);
                q3i[pos] = r.z;
            }
            t3 += __popc(bal);
            __syncwarp();
            if (t3 - h3 >= 32u) {
                const uint4 r3 = q3[(h3 + lane) % QCAP];
                const uint32_t o3 = q3i[(h3 + lane) % QCAP];
                h3 += 32u;
                __syncwarp();
                const Planes p3 = {r3.x, r3.y, r3.z, r3.w};
                stage_b(a, p3, o3, true, tm, lane, feat, what2);
            }
        }
    };
    // ---- stage A2 on up to 32 queued survivors of the HCP test ----------------------------
    auto run_a2 = [&](unsigned count) {
        const uint32_t off = q1[(h1 + lane) % SF_QCAP1];      // index - a.first
        h1 += count;
        __syncwarp();
        const bool active = lane < count;
        const unsigned long long idx = a.first + off;
        uint32_t m0, m1;
        bool ok = active;
        if (EXACT) sf_exact_hand(a.key, a.cons.x, s_low, active ? idx : 0ull, active, m0, m1);
        else {
            sf_primary_hand(a.key, a.cons.sf_tab, a.cons.sf_bmul, s_low, active ? idx : 0ull, active, m0, m1);
            ok = ok && seat_ranges_ok(c, hand_features(m0, m1));
        }
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
                q1[(t1 + __popc(bal & lt)) % SF_QCAP1] = 4u * (co0 + (b >> 2) * stride) + (b & 3u) - mis;
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
        const uint32_t o3 = q3i[(h3 + lane) % QCAP];
        const Planes p3 = {r3.x, r3.y, r3.z, r3.w};
        stage_b(a, p3, o3, lane < t3 - h3, tm, lane, feat, what2);
    }

    {
        unsigned long long v64 = n_void;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v64 += __shfl_xor_sync(0xFFFFFFFFu, v64, o);
        if (lane == 0 && v64) atomicAdd(&a.tallies[1], v64);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&a.tallies[0], a.n);   // the grid covers exactly [first, first + n)
    }
    __syncthreads();
    flush_hist(a, sp, smem, SF_HIST_COPIES);
}


#ifndef BMC_JIT
template <bool HAS_CONS, bool FIXED>
__global__ void __launch_bounds__(THREADS) sample_kernel(const __grid_constant__ SampleArgs a)
{
    sample_body<HAS_CONS, FIXED>(a);
}
template <bool EXACT>
__global__ void __launch_bounds__(THREADS, SF_MIN_BLOCKS) sample_seatfirst_kernel(const __grid_constant__ SampleArgs a)
{
    sample_seatfirst_body<EXACT>(a);
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
#ifdef BMC_JIT_EXACT
    sample_seatfirst_body<true>(a);
#else
    sample_seatfirst_body<false>(a);
#endif
}
#endif
#endif
