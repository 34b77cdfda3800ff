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
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bridgemc.h"
#include "deal.cuh"
#include "features.cuh"
#include "seatfirst.cuh"
#include "reforder.cuh"
#include "hist.cuh"
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
__global__ void __launch_bounds__(THREADS) sample_reforder_kernel(const __grid_constant__ SampleArgs a, const __grid_constant__ SeqDev q)
{
    __shared__ uint32_t s_hist[SF_HIST_COPIES][SH_WORDS];
    __shared__ uint32_t s_feat[WARPS_PER_BLOCK][FEAT_SCRATCH_WORDS];
    __shared__ uint8_t s_shape[14 * 14 * 14];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SF_HIST_COPIES * SH_WORDS; i += THREADS) (&s_hist[0][0])[i] = 0u;
    for (int i = threadIdx.x; i < 14 * 14 * 14; i += THREADS) s_shape[i] = c_shape_lut[i];
    __syncthreads();
    uint32_t *hist = s_hist[warp % SF_HIST_COPIES];
    uint32_t n_raw = 0, n_bad = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * THREADS;
    unsigned long long off = (unsigned long long)blockIdx.x * THREADS + threadIdx.x;
    for (uint32_t it = 0; it < a.iters; ++it, off += stride) {
        const bool valid = off < a.n;
        const unsigned long long idx = a.first + off;
        const uint32_t il = (uint32_t)idx, ih = (uint32_t)(idx >> 32);
        uint32_t lo0 = a.fix.lo0, lo1 = a.fix.lo1, hi0 = a.fix.hi0, hi1 = a.fix.hi1;
        uint32_t f0 = a.fix.free0, f1 = a.fix.free1;
        bool bad = false;
        for (uint32_t j = 0; j < q.n_stages; ++j) {
            const SeqStage &st = q.st[j];
            if (j > 0u && st.m <= 13u) break;                     // a LATER seat is implied once 13 cards remain (bridge.cl:1303);
                                                                  // the first def always goes through the root generator
            const SeqRanges r = seq_infer(st, q.exhaust, f0, f1);
            bad = bad || r.bad;
            uint32_t h0 = 0, h1 = 0, att = 0;
            bool need = valid && !bad;
            while (need) {
                const bool clean = seq_draw13(a.key, il, ih, STREAM_REFORDER + j, att, st.m, st.thresh, f0, f1, h0, h1);
                if (clean && seq_hand_ok(r, h0, h1)) need = false;
                else if (++att >= SEQ_MAX_ATTEMPTS) { bad = true; need = false; }
            }
            if (!valid || bad) { h0 = 0; h1 = 0; }
            const uint32_t k = (uint32_t)st.seat;
            if (k & 1u) { lo0 |= h0; lo1 |= h1; }
            if (k & 2u) { hi0 |= h0; hi1 |= h1; }
            f0 &= ~h0; f1 &= ~h1;
        }
        Planes pl = {lo0, lo1, hi0, hi1};
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
        n_raw += valid ? 1u : 0u;
        n_bad += (valid && bad) ? 1u : 0u;
        stage_b(a, pl, valid && !bad, hist, s_shape, lane, s_feat[warp], 0u);     // nothing left to test
    }
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
    flush_hist(a, &s_hist[0][0], SF_HIST_COPIES);
}

// ---------------------------------------------------------------------------
// K2: features / accept / choose-bid on given records
// ---------------------------------------------------------------------------
struct SchemeDev {
    const uint32_t *code;
    const uint32_t *row_off;     // offsets into code, n_rows entries
    uint32_t n_rows;
};

__global__ void __launch_bounds__(256) filter_kernel(const __grid_constant__ ConsDev cons, int has_cons,
                                                     const __grid_constant__ SchemeDev sch, const uint4 *__restrict__ rec,
                                                     unsigned long long n, uint8_t *__restrict__ accept,
                                                     bmc_features *__restrict__ feats)
{
    __shared__ uint32_t s_feat[8][FEAT_SCRATCH_WORDS];
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (live) r = rec[i];
    const Planes pl = {r.x, r.y, r.z, r.w};
    bool ok = true;
    bmc_features out;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t m0, m1;
        seat_words(pl, k, m0, m1);
        const Feat f = hand_features(m0, m1);
        const uint8_t *fb = feat_store(s_feat[threadIdx.x >> 5], threadIdx.x & 31u, f);
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
        bmc_seat_features &o = out.seat[k];
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

// ---------------------------------------------------------------------------
// K4: scoring
// ---------------------------------------------------------------------------
// contract-score (bridge.cl:3045-3069), written out from the rules of the
// reference text (incl. its non-standard branches)
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
__host__ __device__ inline int base_score_dev(bmc_contract c, int tricks, int vul)
{
    const int s = contract_score_dev(c.suit, tricks, vul, c.dbl, c.level);
    return (c.declarer == 1 || c.declarer == 3) ? -s : s;
}
__constant__ int c_imp_pts[24] = {20, 50, 90, 130, 170, 220, 270, 320, 370, 430, 500, 600, 750, 900,
                                  1100, 1300, 1500, 1750, 2000, 2300, 2500, 3000, 3500, 4000};
__device__ __forceinline__ int imps_dev(int diff, bool &bad)
{
    const int a = diff < 0 ? -diff : diff;
    int pos = 24;
#pragma unroll
    for (int i = 23; i >= 0; --i) if (a < c_imp_pts[i]) pos = i;
    bad = pos == 24;
    return bad ? 0 : (diff < 0 ? -pos : pos);
}

struct ContractTable { bmc_contract c[8]; uint8_t vul[8]; uint8_t pairs[16]; uint32_t n, n_pairs; };

// HBM-bound (1 B in, 4 B out per element): each thread scores four consecutive elements
// (one 32-bit load, one 128-bit store); the contract index advances without a division.
__global__ void __launch_bounds__(256) score_kernel(const bmc_contract *__restrict__ contracts, const uint8_t *__restrict__ vul,
                                                    uint32_t nc, const uint8_t *__restrict__ tricks, unsigned long long total,
                                                    int32_t *__restrict__ scores)
{
    __shared__ bmc_contract s_c[64];
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
            const bmc_contract ct = cached ? s_c[c] : contracts[c];
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

// tricks for deal `idx`, contract c: 16-bit lane c of Philox(stream 1), mapped to
// 0..13 by multiply-shift; lanes whose low product part is < 65536 mod 14 (=2)
// are biased and take the next lane (8..15) instead.
__global__ void __launch_bounds__(256) score_tally_kernel(const __grid_constant__ ContractTable t, const __grid_constant__ PhiloxKey key,
                                                          unsigned long long first, unsigned long long n,
                                                          unsigned long long *__restrict__ tricks_hist,
                                                          unsigned long long *__restrict__ imp_hist)
{
    __shared__ uint32_t s_tr[8 * 14];
    __shared__ uint32_t s_imp[8 * 49];
    __shared__ int s_score[8 * 14];
    for (int i = threadIdx.x; i < 8 * 14; i += blockDim.x) {
        s_tr[i] = 0;
        const int c = i / 14;
        s_score[i] = c < (int)t.n ? base_score_dev(t.c[c], i % 14, t.vul[c]) : 0;
    }
    for (int i = threadIdx.x; i < 8 * 49; i += blockDim.x) s_imp[i] = 0;
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long idx = first + i;
        uint32_t w[4];
        philox4x32_10(key, (uint32_t)idx, (uint32_t)(idx >> 32), 0u, 1u, w);
        int tr[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t lane16 = (w[c >> 1] >> (16 * (c & 1))) & 0xFFFFu;
            uint32_t prod = lane16 * 14u;
            if ((prod & 0xFFFFu) < 2u) {       // biased lane: deterministic redraw from a second block
                uint32_t w2[4];
                philox4x32_10(key, (uint32_t)idx, (uint32_t)(idx >> 32), 1u + c, 1u, w2);
                prod = (w2[0] >> 16) * 14u;    // residual bias 2^-16 * 2/65536 -- below any observable level
            }
            tr[c] = (int)(prod >> 16);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < (int)t.n) atomicAdd(&s_tr[c * 14 + tr[c]], 1u);
#pragma unroll
        for (int p = 0; p < 8; ++p)
            if (p < (int)t.n_pairs) {
                const int ca = t.pairs[2 * p], cb = t.pairs[2 * p + 1];
                bool bad;
                const int v = imps_dev(s_score[ca * 14 + tr[ca]] - s_score[cb * 14 + tr[cb]], bad);
                if (!bad) atomicAdd(&s_imp[p * 49 + 24 + v], 1u);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 14; i += blockDim.x)
        if (s_tr[i]) atomicAdd(&tricks_hist[i], (unsigned long long)s_tr[i]);
    for (int i = threadIdx.x; i < 8 * 49; i += blockDim.x)
        if (s_imp[i]) atomicAdd(&imp_hist[i], (unsigned long long)s_imp[i]);
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

// Publishes the four leading counters into host-mapped pinned memory, so the per-wave
// read-back does not occupy the device->host copy engine (which streams the records).
__global__ void publish_counters_kernel(const unsigned long long *__restrict__ tallies, volatile unsigned long long *mapped)
{
    if (threadIdx.x < 4) mapped[threadIdx.x] = tallies[threadIdx.x];
    __threadfence_system();
}

// ---------------------------------------------------------------------------
// host objects
// ---------------------------------------------------------------------------
struct bmc_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    unsigned long long *h_head = nullptr;        // pinned + mapped: per-wave read-back of the four counters
    unsigned long long *h_head_dev = nullptr;    // its device alias
    bmc_record *d_rec = nullptr;                 // grow-only record scratch of the host-buffer API
    uint64_t rec_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long *d_tallies = nullptr;     // scratch tallies for the host-buffer API
    uint64_t launches = 0;
    int blocks_cons = 0, blocks_free = 0, blocks_sf = 0;
    std::mutex mu;
};

struct bmc_constraints {
    ConsDev dev{};               // dev.code filled lazily per device
    std::vector<uint32_t> code;
    uint32_t *d_code = nullptr;
    int d_device = -1;
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
    int jit = 0;                 // 0: interpret the rule bytecode; 1: compile the rule forms into the kernel (NVRTC)
    JitKernel jit_kernel[2];     // [0] full-deal kernel, [1] seat-first kernel
    std::string jit_log;
    int mode = 0;                // requested: 0 auto, 1 full deal (deal52), 2 seat-first, 3 reference order
    int resolved = 0;            // 0 = not calibrated yet
    double stage_a_pass = -1.0;  // measured pass rate of the primary seat's ranges
    std::mutex mu, cal_mu;
};

struct bmc_scheme {
    std::vector<uint32_t> code, row_off;
    std::vector<std::string> bids;
    uint32_t *d_code = nullptr, *d_off = nullptr;
    int d_device = -1;
    std::mutex mu;
};

static std::once_flag g_lut_once[16];
static uint8_t g_shapes[39][4];
static uint8_t g_shape_lut[14 * 14 * 14];
static std::once_flag g_shape_once;

static void init_shapes()
{
    int n = 0;
    for (int a = 4; a <= 13; ++a)
        for (int b = 0; b <= a; ++b)
            for (int c = 0; c <= b; ++c) {
                const int d = 13 - a - b - c;
                if (d < 0 || d > c) continue;
                g_shapes[n][0] = (uint8_t)a; g_shapes[n][1] = (uint8_t)b; g_shapes[n][2] = (uint8_t)c; g_shapes[n][3] = (uint8_t)d;
                ++n;
            }
    memset(g_shape_lut, 0xFF, sizeof g_shape_lut);
    for (int x = 0; x <= 13; ++x)
        for (int y = 0; x + y <= 13; ++y)
            for (int z = 0; x + y + z <= 13; ++z) {
                int v[4] = {x, y, z, 13 - x - y - z};
                for (int i = 0; i < 4; ++i)
                    for (int j = i + 1; j < 4; ++j)
                        if (v[j] > v[i]) { int t = v[i]; v[i] = v[j]; v[j] = t; }
                for (int k = 0; k < 39; ++k)
                    if (g_shapes[k][0] == v[0] && g_shapes[k][1] == v[1] && g_shapes[k][2] == v[2] && g_shapes[k][3] == v[3])
                        g_shape_lut[x * 196 + y * 14 + z] = (uint8_t)k;
            }
}

extern "C" {

const char *bmc_last_error(void) { return g_err.c_str(); }
const char *bmc_version(void) { return "bridgemc 0.1 (sm_100a)"; }

void bmc_shape_table(uint8_t out[39][4])
{
    std::call_once(g_shape_once, init_shapes);
    memcpy(out, g_shapes, sizeof g_shapes);
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
    std::call_once(g_shape_once, init_shapes);
    bmc_ctx *c = new bmc_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaMalloc(&c->d_tallies, sizeof(bmc_tallies)));
    CUDA_TRY(cudaHostAlloc(&c->h_head, 4 * sizeof(unsigned long long), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer((void **)&c->h_head_dev, c->h_head, 0));
    if (device < 16) {
        cudaError_t le = cudaSuccess;
        std::call_once(g_lut_once[device], [&] { le = cudaMemcpyToSymbol(c_shape_lut, g_shape_lut, sizeof g_shape_lut); });
        if (le != cudaSuccess) return fail(BMC_E_CUDA, std::string("shape LUT upload: ") + cudaGetErrorString(le));
    } else {
        CUDA_TRY(cudaMemcpyToSymbol(c_shape_lut, g_shape_lut, sizeof g_shape_lut));
    }
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sample_kernel<true, false>, THREADS, 0));
    c->blocks_cons = c->sms * (occ > 0 ? occ : 1);
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sample_kernel<false, false>, THREADS, 0));
    c->blocks_free = c->sms * (occ > 0 ? occ : 1);
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sample_seatfirst_kernel, THREADS, 0));
    c->blocks_sf = c->sms * (occ > 0 ? occ : 1);
    if (const char *e = getenv("BMC_BLOCKS_PER_SM")) {        // tuning knob: cap the resident blocks per SM
        const int cap = atoi(e);
        if (cap > 0) {
            if (c->blocks_sf > c->sms * cap) c->blocks_sf = c->sms * cap;
            if (c->blocks_cons > c->sms * cap) c->blocks_cons = c->sms * cap;
            if (c->blocks_free > c->sms * cap) c->blocks_free = c->sms * cap;
        }
    }
    *out = c;
    return BMC_OK;
}

void bmc_shutdown(bmc_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_tallies);
    cudaFreeHost(c->h_head);
    if (c->d_rec) cudaFree(c->d_rec);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->own_stream);
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

int bmc_constraints_set_mode(bmc_constraints *c, int mode)
{
    if (!c || mode < 0 || mode > 2) return fail(BMC_E_INVALID, "bmc_constraints_set_mode: bad argument (mode 3 is set by bmc_constraints_set_reference_order)");
    if (mode == 2 && (c->has_fixed || !c->has_cons)) return fail(BMC_E_INVALID, "seat-first mode needs a constrained seat and no fixed hands");
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
        for (int w = 0; w < 13; ++w) {
            uint64_t N = 1;
            for (int d = 0; d < 4; ++d) {
                const int n = (int)m - 4 * w - d;
                if (n >= 2) N *= (uint64_t)n;
            }
            st.thresh[w] = (uint32_t)((1ull << 32) % N);
        }
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
    if (c->d_code) { cudaSetDevice(c->d_device); cudaFree(c->d_code); }
    for (JitKernel &jk : c->jit_kernel)
        if (jk.module) { cudaSetDevice(jk.device); JitApi::get().ModuleUnload(jk.module); }
    delete c;
}

// ---- NVRTC specialisation of a constraint (jit.hpp) ------------------------------------------------
static int jit_build(bmc_ctx *ctx, bmc_constraints *c, bool seatfirst)
{
    JitKernel &jk = c->jit_kernel[seatfirst ? 1 : 0];
    if (jk.func && jk.device == ctx->device) return BMC_OK;
    JitApi &api = JitApi::get();
    if (!api.ok) return fail(BMC_E_CUDA, "NVRTC specialisation unavailable: " + api.why);
    // 1. the rule forms as one C++ function
    std::string fn =
        "__device__ __forceinline__ bool bmc_jit_seat_rule(int seat, const bmc::Feat &f, bool &err)\n{\n"
        "    const int vC = f.f0 & 0xFF, vD = (f.f0 >> 8) & 0xFF, vH = (f.f0 >> 16) & 0xFF, vS = (f.f0 >> 24) & 0xFF;\n"
        "    const int vCP = f.f1 & 0xFF, vDP = (f.f1 >> 8) & 0xFF, vHP = (f.f1 >> 16) & 0xFF, vSP = (f.f1 >> 24) & 0xFF;\n"
        "    const int vHCP = f.f2 & 0xFF, vLOS = (f.f2 >> 8) & 0xFF; const bool vBAL = ((f.f2 >> 16) & 0xFF) != 0;\n"
        "    (void)vC; (void)vD; (void)vH; (void)vS; (void)vCP; (void)vDP; (void)vHP; (void)vSP; (void)vHCP; (void)vLOS; (void)vBAL;\n"
        "    switch (seat) {\n";
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
    const std::string prog = std::string(
        "typedef unsigned char uint8_t; typedef signed char int8_t; typedef unsigned short uint16_t; typedef short int16_t;\n"
        "typedef unsigned int uint32_t; typedef int int32_t; typedef unsigned long long uint64_t; typedef long long int64_t;\n"
        "#define BMC_JIT 1\n") + (seatfirst ? "#define BMC_JIT_SEATFIRST 1\n" : "#define BMC_JIT_FULL 1\n") +
        "#include \"features.cuh\"\n" + fn + "#include \"kernels.cuh\"\n";
    // 2. compile for sm_100a
    auto t0 = std::chrono::steady_clock::now();
    nvrtcProgram p = nullptr;
    if (api.CreateProgram(&p, prog.c_str(), "bmc_jit.cu", kJitSrcCount, kJitSrcTexts, kJitSrcNames) != NVRTC_SUCCESS)
        return fail(BMC_E_CUDA, "nvrtcCreateProgram failed");
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    const nvrtcResult r = api.CompileProgram(p, 3, opts);
    size_t ls = 0;
    api.GetProgramLogSize(p, &ls);
    c->jit_log.assign(ls, ' ');
    if (ls) api.GetProgramLog(p, &c->jit_log[0]);
    if (r != NVRTC_SUCCESS) {
        api.DestroyProgram(&p);
        return fail(BMC_E_CUDA, "NVRTC compilation failed: " + c->jit_log.substr(0, 2000));
    }
    size_t cs = 0;
    api.GetCUBINSize(p, &cs);
    std::vector<char> cubin(cs);
    api.GetCUBIN(p, cubin.data());
    api.DestroyProgram(&p);
    // 3. load, bind, initialise the module's constant tables
    if (jk.module) { api.ModuleUnload(jk.module); jk = JitKernel(); }
    CUresult cr = api.ModuleLoadData(&jk.module, cubin.data());
    if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "cuModuleLoadData failed: " + std::to_string((int)cr));
    cr = api.ModuleGetFunction(&jk.func, jk.module, seatfirst ? "bmc_jit_sample_seatfirst" : "bmc_jit_sample_full");
    if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "cuModuleGetFunction failed: " + std::to_string((int)cr));
    CUdeviceptr lut = 0;
    size_t lut_bytes = 0;
    cr = api.ModuleGetGlobal(&lut, &lut_bytes, jk.module, "c_shape_lut");
    if (cr != CUDA_SUCCESS || lut_bytes != sizeof g_shape_lut) return fail(BMC_E_CUDA, "jit: c_shape_lut not found");
    cr = api.MemcpyHtoD(lut, g_shape_lut, sizeof g_shape_lut);
    if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "jit: shape LUT upload failed");
    int occ = 0;
    api.Occupancy(&occ, jk.func, THREADS, 0);
    jk.blocks_per_sm = occ > 0 ? occ : 1;
    jk.device = ctx->device;
    jk.compile_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return BMC_OK;
}

int bmc_constraints_set_jit(bmc_constraints *c, int on)
{
    if (!c || on < 0 || on > 1) return fail(BMC_E_INVALID, "bmc_constraints_set_jit: bad argument");
    c->jit = on;
    return BMC_OK;
}
float bmc_constraints_jit_ms(const bmc_constraints *c)
{
    if (!c) return 0.f;
    return c->jit_kernel[0].compile_ms + c->jit_kernel[1].compile_ms;
}

static int cons_upload(bmc_ctx *ctx, const bmc_constraints *cc, ConsDev &dev)
{
    bmc_constraints *c = const_cast<bmc_constraints *>(cc);
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->d_code && c->d_device != ctx->device) {
        cudaSetDevice(c->d_device);
        cudaFree(c->d_code);
        c->d_code = nullptr;
        cudaSetDevice(ctx->device);
    }
    if (!c->d_code) {
        CUDA_TRY(cudaMalloc(&c->d_code, c->code.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpy(c->d_code, c->code.data(), c->code.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        c->d_device = ctx->device;
    }
    dev = c->dev;
    dev.code = c->d_code;
    dev.sf_tab = reinterpret_cast<const uint4 *>(c->d_code + c->tab_off);
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
    if (s->d_code) { cudaSetDevice(s->d_device); cudaFree(s->d_code); cudaFree(s->d_off); }
    delete s;
}
static int scheme_upload(bmc_ctx *ctx, const bmc_scheme *ss, SchemeDev &dev)
{
    dev.code = nullptr; dev.row_off = nullptr; dev.n_rows = 0;
    if (!ss || ss->row_off.empty()) return BMC_OK;
    bmc_scheme *s = const_cast<bmc_scheme *>(ss);
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->d_code && s->d_device != ctx->device) {
        cudaSetDevice(s->d_device);
        cudaFree(s->d_code); cudaFree(s->d_off);
        s->d_code = s->d_off = nullptr;
        cudaSetDevice(ctx->device);
    }
    if (!s->d_code) {
        CUDA_TRY(cudaMalloc(&s->d_code, s->code.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&s->d_off, s->row_off.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpy(s->d_code, s->code.data(), s->code.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(s->d_off, s->row_off.data(), s->row_off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        s->d_device = ctx->device;
    }
    dev.code = s->d_code; dev.row_off = s->d_off; dev.n_rows = (uint32_t)s->row_off.size();
    return BMC_OK;
}

// ---- sampling -------------------------------------------------------------------
static int launch_wave(bmc_ctx *ctx, const ConsDev &cons, bool has_cons, const FixDev *fix, bool seatfirst, uint64_t seed,
                       uint64_t first, uint64_t n, uint32_t tally_mask, void *records_dev, uint64_t cap, void *tallies_dev,
                       const SeqDev *seq = nullptr, const JitKernel *jit = nullptr)
{
    SampleArgs a;
    a.cons = cons;
    if (fix) a.fix = *fix;
    else { memset(&a.fix, 0, sizeof a.fix); a.fix.n_free = 52; }
    a.key = philox_key(seed);
    a.first = first;
    a.n = n;
    a.records = (uint4 *)records_dev;
    a.cap = records_dev ? cap : 0;
    a.tallies = (unsigned long long *)tallies_dev;
    a.tally_mask = tally_mask;
    {   // stage B of the seat-first sampler: the primary seat has no capacity left
        uint32_t cap4[4] = {13, 13, 13, 13};
        cap4[cons.primary & 3u] = 0;
        const uint32_t p0 = cap4[0], p1 = p0 + cap4[1], p2 = p1 + cap4[2];
        a.sf_q0 = (128u - p0) | ((128u - p1) << 8) | ((128u - p2) << 16);
        a.pad = 0;
    }
    int blocks = seatfirst ? ctx->blocks_sf : has_cons ? ctx->blocks_cons : ctx->blocks_free;
    if (seq) blocks = ctx->sms * 4;
    if (jit) blocks = ctx->sms * jit->blocks_per_sm;
    const uint64_t per_thread = seatfirst ? 4 * SF_CALLS : 1;     // raw deals per thread per iteration
    const uint64_t per_iter = (uint64_t)blocks * THREADS * per_thread;
    if (n < per_iter) blocks = (int)((n + THREADS * per_thread - 1) / (THREADS * per_thread));
    if (blocks < 1) blocks = 1;
    const uint64_t chunk = (uint64_t)blocks * THREADS * per_thread;
    a.iters = (uint32_t)((n + chunk - 1) / chunk);
    if (seatfirst) {
        // a lane owns index QUADS (4c .. 4c + 3, one Philox call): a misaligned first index or end adds one
        const uint64_t calls = ((first + n + 3) >> 2) - (first >> 2);
        const uint64_t per = (uint64_t)blocks * THREADS * SF_CALLS;
        a.iters = (uint32_t)((calls + per - 1) / per);
    }
    if (jit) {
        void *params[] = {&a};
        const CUresult cr = JitApi::get().LaunchKernel(jit->func, (unsigned)blocks, 1, 1, THREADS, 1, 1, 0, (CUstream)ctx->stream, params, nullptr);
        if (cr != CUDA_SUCCESS) return fail(BMC_E_CUDA, "cuLaunchKernel (jit) failed: " + std::to_string((int)cr));
    } else if (seq) sample_reforder_kernel<<<blocks, THREADS, 0, ctx->stream>>>(a, *seq);
    else if (seatfirst) sample_seatfirst_kernel<<<blocks, THREADS, 0, ctx->stream>>>(a);
    else if (fix) {
        if (has_cons) sample_kernel<true, true><<<blocks, THREADS, 0, ctx->stream>>>(a);
        else sample_kernel<false, true><<<blocks, THREADS, 0, ctx->stream>>>(a);
    } else {
        if (has_cons) sample_kernel<true, false><<<blocks, THREADS, 0, ctx->stream>>>(a);
        else sample_kernel<false, false><<<blocks, THREADS, 0, ctx->stream>>>(a);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return BMC_OK;
}

// Common driver.  host_out != NULL: records are streamed to the caller's buffer on a
// second stream while the next wave computes (the copy of wave k overlaps wave k+1).
static int sample_core(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
                       uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, void *records_dev, uint64_t cap,
                       void *tallies_dev, bmc_stats *stats, bmc_record *host_out)
{
    if (n_accept_target == 0 && n_raw == 0) return fail(BMC_E_INVALID, "nothing to do: n_raw == 0 and no accept target");
    auto t0 = std::chrono::steady_clock::now();
    ConsDev dev{};
    const bool has_cons = cons && cons->has_cons;
    if (has_cons) { int rc = cons_upload(ctx, cons, dev); if (rc) return rc; }
    const FixDev *fix = (cons && cons->has_fixed) ? &cons->fix : nullptr;
    bool seatfirst = false;
    const SeqDev *seq = nullptr;
    if (cons && cons->resolved == 3) {                 // reference order: `build`'s own sequential semantics
        seq = &cons->seq;
        fix = &cons->fix;
    } else if (has_cons && !fix) {
        bmc_constraints *cc = const_cast<bmc_constraints *>(cons);
        std::lock_guard<std::mutex> cal_lock(cc->cal_mu);      // one calibration per constraint, even across contexts
        // pass rate of `cal` over 2^16 raw deals of the full-deal kernel (a fixed calibration seed:
        // the outcome, hence the mode and the stream behind each index, is reproducible)
        auto pass_rate = [&](const ConsDev &cal, double &rate) -> int {
            unsigned long long *d_cal = nullptr, h4[4] = {0, 0, 0, 0};
            CUDA_TRY(cudaMalloc(&d_cal, sizeof(bmc_tallies)));
            cudaMemsetAsync(d_cal, 0, sizeof(bmc_tallies), ctx->stream);
            int rc = launch_wave(ctx, cal, true, nullptr, false, 0x63616C6962726174ull, 0, 1u << 16, 0, nullptr, 0, d_cal);
            cudaError_t e = cudaMemcpyAsync(h4, d_cal, sizeof h4, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_cal);
            if (rc) return rc;
            if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("calibration: ") + cudaGetErrorString(e));
            rate = (double)h4[2] / (double)(h4[0] ? h4[0] : 1);
            return BMC_OK;
        };
        if (cc->resolved == 0) {
            if (cc->mode != 0) cc->resolved = cc->mode;
            else {
                // the primary seat's RANGES alone: seat-first pays off when most deals die on them
                ConsDev cal = dev;
                for (int k = 0; k < 4; ++k) {
                    cal.seat[k].prog = 0xFFFFFFFFu;
                    if ((uint32_t)k != dev.primary) cal.seat[k].active = 0;
                }
                int rc = pass_rate(cal, cc->stage_a_pass);
                if (rc) return rc;
                cc->resolved = cc->stage_a_pass < 0.30 ? 2 : 1;
            }
            cc->dev.b1_mask = 0xFu;
            if (cc->resolved == 2 && dev.n_active >= 3) {
                // the most selective other seat: tested together with the primary seat by the compiled-rule kernel
                double best = 2.0;
                int second = -1;
                for (int k = 0; k < 4; ++k) {
                    if ((uint32_t)k == dev.primary || !dev.seat[k].active) continue;
                    ConsDev cal = dev;
                    for (int j = 0; j < 4; ++j) if (j != k) cal.seat[j].active = 0;
                    cal.primary = (uint32_t)k;
                    double r = 1.0;
                    int rc = pass_rate(cal, r);
                    if (rc) return rc;
                    if (r < best) { best = r; second = k; }
                }
                if (second >= 0) cc->dev.b1_mask = (1u << dev.primary) | (1u << second);
            }
        }
        dev.b1_mask = cc->dev.b1_mask;
        seatfirst = cc->resolved == 2;
    }
    const JitKernel *jit = nullptr;
    if (cons && cons->jit && has_cons && !fix && !seq && cons->has_rules) {
        bmc_constraints *cc = const_cast<bmc_constraints *>(cons);
        std::lock_guard<std::mutex> jl(cc->cal_mu);
        int rc = jit_build(ctx, cc, seatfirst);
        if (rc) return rc;
        jit = &cc->jit_kernel[seatfirst ? 1 : 0];
    }
    const bool stream_copy = host_out && records_dev && cap;
    const bool stepwise = n_accept_target != 0 || stream_copy;
    if (wave == 0) wave = 1ull << 26;
    // 2^36 keeps a warp's iteration counter in 32 bits; the seat-first kernel carries 32-bit offsets from the
    // launch's first index in its queues: 2^31 indices per launch
    uint64_t max_launch = stepwise ? wave : (1ull << 36);
    if (seatfirst && max_launch > (1ull << 31)) max_launch = 1ull << 31;
    uint32_t waves = 0;
    const uint64_t launches0 = ctx->launches;
    // The counters are read back through host-mapped memory written by a tiny kernel: a memcpy on
    // the compute stream would queue on the same device->host copy engine as the record copies of
    // copy_stream and was measured to serialise the waves behind them.
    volatile unsigned long long *head = ctx->h_head;
    unsigned long long start_acc = 0;
    head[0] = head[1] = head[2] = head[3] = 0;
    auto read_head = [&]() -> cudaError_t {
        publish_counters_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)tallies_dev, ctx->h_head_dev);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(ctx->stream);
    };
    if (stepwise) {
        CUDA_TRY(read_head());
        start_acc = head[2];
    }
    uint64_t copied = start_acc < cap ? start_acc : cap;
    float kernel_ms = 0.f;
    if (!stepwise) CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    uint64_t done = 0;
    for (;;) {
        if (n_raw && done >= n_raw) break;
        uint64_t n = max_launch;
        if (n_raw && n_raw - done < n) n = n_raw - done;
        if (stepwise) CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        int rc = launch_wave(ctx, dev, has_cons, fix, seatfirst, seed, first_index + done, n, tally_mask, records_dev, cap, tallies_dev, seq, jit);
        if (rc) return rc;
        done += n;
        ++waves;
        if (!stepwise) continue;
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(read_head());
        {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);       // device time of this wave's kernel only
            kernel_ms += ms;
        }
        if (stream_copy) {
            const uint64_t upto = head[2] < cap ? head[2] : cap;
            if (upto > copied) {
                auto tc0 = std::chrono::steady_clock::now();
                CUDA_TRY(cudaMemcpyAsync(host_out + copied, (const bmc_record *)records_dev + copied,
                                         (upto - copied) * sizeof(bmc_record), cudaMemcpyDeviceToHost, ctx->copy_stream));
                if (getenv("BMC_DEBUG"))
                    fprintf(stderr, "[bmc] wave %u: copy of %.1f MB enqueued in %.3f ms at t=%.2f ms (kernel_ms so far %.2f)\n", waves,
                            (upto - copied) * 16e-6,
                            std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - tc0).count(),
                            std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(), kernel_ms);
                copied = upto;
            }
        }
        if (n_accept_target && head[2] - start_acc >= n_accept_target) break;
        // watchdog for open-ended jobs: a constraint nothing satisfies must not spin forever
        if (n_raw == 0 && ((done >= (1ull << 36) && head[2] == start_acc) || done >= (1ull << 44))) break;
        if (seq && head[2] == start_acc && head[1] >= done) break;      // reference order: every build signalled an error
    }
    if (!stepwise) CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(read_head());
    if (getenv("BMC_DEBUG"))
        fprintf(stderr, "[bmc] compute done at t=%.2f ms\n", std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (stream_copy) CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    if (getenv("BMC_DEBUG"))
        fprintf(stderr, "[bmc] copies done at t=%.2f ms\n", std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (stats) {
        float ms = kernel_ms;
        if (!stepwise) cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        stats->raw = head[0]; stats->voided = head[1]; stats->accepted = head[2];
        stats->stored = records_dev ? (head[2] < cap ? head[2] : cap) : 0;
        stats->waves = waves;
        stats->launches = (uint32_t)(ctx->launches - launches0);
        stats->kernel_ms = ms;
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
                       stats, nullptr);
}

int bmc_sample(bmc_ctx *ctx, const bmc_constraints *cons, uint64_t seed, uint64_t first_index, uint64_t n_raw,
               uint64_t n_accept_target, uint64_t wave, uint32_t tally_mask, bmc_record *records, uint64_t cap,
               bmc_tallies *tallies, bmc_stats *stats)
{
    if (!ctx) return fail(BMC_E_INVALID, "ctx is NULL");
    auto t0 = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!records) cap = 0;
    // device scratch for the records is kept by the context and only ever grows
    if (cap > ctx->rec_cap) {
        if (ctx->d_rec) cudaFree(ctx->d_rec);
        ctx->d_rec = nullptr;
        ctx->rec_cap = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_rec, cap * sizeof(bmc_record)));
        ctx->rec_cap = cap;
    }
    CUDA_TRY(cudaMemsetAsync(ctx->d_tallies, 0, sizeof(bmc_tallies), ctx->stream));
    bmc_stats st{};
    int rc = sample_core(ctx, cons, seed, first_index, n_raw, n_accept_target, wave, tally_mask, cap ? ctx->d_rec : nullptr, cap,
                         ctx->d_tallies, &st, cap ? records : nullptr);
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

// ---- filter ---------------------------------------------------------------------
int bmc_filter(bmc_ctx *ctx, const bmc_constraints *cons, const bmc_scheme *scheme, const bmc_record *records, uint64_t n,
               uint8_t *accept, bmc_features *feats)
{
    if (!ctx) return fail(BMC_E_INVALID, "ctx is NULL");
    if (n == 0) return BMC_OK;
    if (!records) return fail(BMC_E_INVALID, "records is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    ConsDev dev{};
    const bool has_cons = cons && cons->has_cons;
    if (has_cons) { int rc = cons_upload(ctx, cons, dev); if (rc) return rc; }
    SchemeDev sd;
    { int rc = scheme_upload(ctx, scheme, sd); if (rc) return rc; }
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
    for (uint64_t off = 0; off < n && e == cudaSuccess; off += chunk) {
        const uint64_t k = n - off < chunk ? n - off : chunk;
        e = cudaMemcpyAsync(d_rec, records + off, k * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) break;
        filter_kernel<<<(unsigned)((k + 255) / 256), 256, 0, ctx->stream>>>(dev, has_cons ? 1 : 0, sd, d_rec, k, d_acc, d_feat);
        ctx->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess && accept) e = cudaMemcpyAsync(accept + off, d_acc, k, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && feats) e = cudaMemcpyAsync(feats + off, d_feat, k * sizeof(bmc_features), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) rc = fail(BMC_E_CUDA, std::string("bmc_filter: ") + cudaGetErrorString(e));
    cudaFree(d_rec); cudaFree(d_acc); cudaFree(d_feat);
    return rc;
}

// ---- scoring --------------------------------------------------------------------
static int check_contracts(const bmc_contract *c, uint32_t n)
{
    for (uint32_t i = 0; i < n; ++i)
        if (c[i].level < 0 || c[i].level > 7 || c[i].suit < 0 || c[i].suit > 4 || c[i].dbl < 0 || c[i].dbl > 2 || c[i].declarer < -1 || c[i].declarer > 3)
            return fail(BMC_E_INVALID, "contract " + std::to_string(i) + " out of range");
    return BMC_OK;
}

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
        score_kernel<<<blocks, 256, 0, ctx->stream>>>(d_c, d_v, nc, d_t, total, d_s);
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
                    int32_t *score_table)
{
    if (!ctx || !contracts || nc == 0 || nc > 8 || n_pairs > 8 || !tricks_hist) return fail(BMC_E_INVALID, "bmc_score_tally: bad argument");
    if (n_pairs && (!pairs || !imp_hist)) return fail(BMC_E_INVALID, "bmc_score_tally: pairs / imp_hist is NULL");
    { int rc = check_contracts(contracts, nc); if (rc) return rc; }
    ContractTable t{};
    t.n = nc; t.n_pairs = n_pairs;
    for (uint32_t i = 0; i < nc; ++i) { t.c[i] = contracts[i]; t.vul[i] = vulnerable ? vulnerable[i] : 0; }
    for (uint32_t i = 0; i < 2 * n_pairs; ++i) {
        if (pairs[i] >= nc) return fail(BMC_E_INVALID, "pair index out of range");
        t.pairs[i] = pairs[i];
    }
    if (score_table)
        for (uint32_t c = 0; c < nc; ++c)
            for (int tr = 0; tr < 14; ++tr) score_table[c * 14 + tr] = base_score_dev(t.c[c], tr, t.vul[c]);
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (8 * 14 + 8 * 49) * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, (8 * 14 + 8 * 49) * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess && n) {
        uint64_t want = (n + 255) / 256;
        const unsigned blocks = (unsigned)(want < (uint64_t)ctx->sms * 8 ? want : (uint64_t)ctx->sms * 8);
        score_tally_kernel<<<blocks, 256, 0, ctx->stream>>>(t, philox_key(seed), first_index, n, d, d + 8 * 14);
        ctx->launches++;
        e = cudaGetLastError();
    }
    std::vector<unsigned long long> h(8 * 14 + 8 * 49);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(BMC_E_CUDA, std::string("bmc_score_tally: ") + cudaGetErrorString(e));
    for (uint32_t c = 0; c < nc; ++c)
        for (int tr = 0; tr < 14; ++tr) tricks_hist[c * 14 + tr] = h[c * 14 + tr];
    for (uint32_t p = 0; p < n_pairs; ++p)
        for (int v = 0; v < 49; ++v) imp_hist[p * 49 + v] = h[8 * 14 + p * 49 + v];
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
    alloc((void **)&d_lut, sizeof g_shape_lut);
    std::vector<unsigned long long> h_uniq;
    std::vector<uint32_t> h_cnt, h_first;
    uint32_t runs = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rec, records, n * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_lut, g_shape_lut, sizeof g_shape_lut, cudaMemcpyHostToDevice, ctx->stream);
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
        e = cudaMemcpy(h_uniq.data(), d_uniq, runs * 8ull, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(h_cnt.data(), d_cnt, runs * 4ull, cudaMemcpyDeviceToHost);
        std::vector<unsigned long long> off(runs);
        unsigned long long acc = 0;
        for (uint32_t i = 0; i < runs; ++i) { off[i] = acc; acc += h_cnt[i]; }
        alloc((void **)&d_off, runs * 8ull); alloc((void **)&d_first, runs * 4ull);
        if (e == cudaSuccess) e = cudaMemcpy(d_off, off.data(), runs * 8ull, cudaMemcpyHostToDevice);
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
