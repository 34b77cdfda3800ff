// features.cuh -- hand features, range test and the rule interpreter (device).
//
// Replaces assess-hand / suit-hcp / suit-loosers / balanced? (bidding.cl:20-46,
// bridge.cl:667-670,2986-2987) and good-opening?'s sublis+eval (bidding.cl:60-74).
// A seat's hand is two 32-bit words of 16-bit suit lanes (m0 = clubs|diamonds<<16,
// m1 = hearts|spades<<16).  All eleven features are produced SWAR, one byte per
// suit:
//     f0 = lens   (bytes c d h s)         f1 = power / suit HCP (bytes c d h s)
//     f2 = hcp | losers<<8 | balanced<<16
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif
#include "bytecode.h"

namespace bmc {

struct Feat { uint32_t f0, f1, f2; };

// Per-seat compiled constraint.  Range test on packed bytes: with
// A = 0x80 - lo and B = 0x7F - hi per byte, v in [lo,hi] for every byte iff
// ((v + A) & ~(v + B) & 0x80808080) == 0x80808080.
struct SeatC {
    uint32_t A[3], B[3];
    uint32_t prog;          // offset of the seat's program in the code buffer, 0xFFFFFFFF = none
    uint32_t active;        // 0: seat unconstrained
    uint32_t need;          // bit 1: some suit-HCP (power) range is not the default
    uint32_t hcp_lo_span;   // hcp_lo | (hcp_hi - hcp_lo) << 8: one subtract + compare in the seat-first stage A1
};

__device__ __forceinline__ Feat hand_features(uint32_t m0, uint32_t m1)
{
    const uint32_t M = 0x01010101u;
    // lengths
    const uint32_t lc = __popc(m0 & 0xFFFFu), ld = __popc(m0 >> 16), lh = __popc(m1 & 0xFFFFu), ls = __popc(m1 >> 16);
    const uint32_t L = lc | (ld << 8) | (lh << 16) | (ls << 24);
    // honour nibbles (J Q K A = lane bits 9..12), one per byte in order c d h s
    const uint32_t x0 = (m0 >> 9) & 0x000F000Fu;          // c at byte 0, d at byte 2
    const uint32_t x1 = (m1 >> 9) & 0x000F000Fu;          // h at byte 0, s at byte 2
    const uint32_t X = __byte_perm(x0, x1, 0x6420);       // bytes: c, d, h, s
    const uint32_t J = X & M, Qn = (X >> 1) & M, K = (X >> 2) & M, A = (X >> 3) & M;
    const uint32_t P = J + 2u * Qn + 3u * K + 4u * A;     // rank-hcp: J1 Q2 K3 A4, <= 10 per byte
    const uint32_t hcp = (P * M) >> 24;
    // losers (bidding.cl:20-23): [no A & len>=1] + [no K & len>=2] + [no Q & len>=3]
    const uint32_t g1 = ((L + 0x7F7F7F7Fu) >> 7) & M;
    const uint32_t g2 = ((L + 0x7E7E7E7Eu) >> 7) & M;
    const uint32_t g3 = ((L + 0x7D7D7D7Du) >> 7) & M;
    const uint32_t lb = (g1 & ~A) + (g2 & ~K) + (g3 & ~Qn);
    const uint32_t losers = (lb * M) >> 24;
    // balanced? (bidding.cl:40-46)
    const bool shape_ok = g2 == M && ((L + 0x7B7B7A7Au) & 0x80808080u) == 0u;   // all >= 2; c,d <= 5; h,s <= 4
    const uint32_t weak = __popc(~(P + 0x7D7D7D7Du) & 0x80808080u);             // suits with power < 3
    const uint32_t bal = (shape_ok && (hcp < 12u || weak < 2u)) ? 1u : 0u;
    Feat f;
    f.f0 = L;
    f.f1 = P;
    f.f2 = hcp | (losers << 8) | (bal << 16);
    return f;
}

__device__ __forceinline__ bool in_ranges(uint32_t v, uint32_t A, uint32_t B)
{
    return (((v + A) & ~(v + B)) & 0x80808080u) == 0x80808080u;
}

__device__ __forceinline__ bool seat_ranges_ok(const SeatC &c, const Feat &f)
{
    return in_ranges(f.f0, c.A[0], c.B[0]) && in_ranges(f.f1, c.A[1], c.B[1]) && in_ranges(f.f2, c.A[2], c.B[2]);
}

// Per-lane scratch for the interpreter: the eleven feature bytes of the hand under test live in
// shared memory (12 bytes per lane, lane stride 3 words: conflict-free), so an operand fetch is
// one LDS.U8 with a warp-uniform offset instead of a register select + shift.
constexpr int FEAT_SCRATCH_WORDS = 32 * 3;     // per warp

__device__ __forceinline__ uint8_t *feat_store(uint32_t *warp_scratch, unsigned lane, const Feat &f)
{
    uint32_t *p = warp_scratch + 3u * lane;
    p[0] = f.f0; p[1] = f.f1; p[2] = f.f2;
    return reinterpret_cast<uint8_t *>(p);
}

// Interpret one program (warp-uniform control flow; the boolean stack lives in one 32-bit register,
// two bits per slot -- value, error -- depth <= 15 checked by the compiler).  `fb` = this lane's
// feature bytes (feat_store).  The next word is fetched while the current one executes.  err is set
// where the reference would signal an "illegal function call" (un-spliced rule lists,
// bidding.cl:249) -- i.e. only when Lisp's left-to-right short-circuit evaluation REACHES the
// illegal form: programs that contain one use the three-valued ops (bytecode.h); invariant: value
// bit = 0 wherever the error bit is set.  Programs without one never set an error bit and keep the
// cheap two-valued ops.
__device__ __forceinline__ bool run_program(const uint32_t *__restrict__ code, const uint8_t *fb, bool &err)
{
    uint32_t st = 0;
    uint32_t w = __ldg(code);
    for (;;) {
        const uint32_t op = w & 15u;
        if (op == OP_END) break;
        const uint32_t next = __ldg(++code);          // every code buffer ends with a spare OP_END
        if (op <= OP_CMPF) {                          // OP_CMPC / OP_CMPF
            const uint32_t a = fb[(w >> 8) & 15u];
            const uint32_t b = op == OP_CMPC ? (w >> 16) & 0xFFu : (uint32_t)fb[(w >> 12) & 15u];
            const uint32_t state = a < b ? 0u : (a == b ? 1u : 2u);
            st = (st << 2) | (((w >> 4) >> state) & 1u);
        } else if (op == OP_AND) { const uint32_t t = st & 1u; st >>= 2; st &= ~1u | t; }
        else if (op == OP_OR) { const uint32_t t = st & 1u; st >>= 2; st |= t; }
        else if (op == OP_NOT) st ^= 1u;
        else if (op == OP_TRUE) st = (st << 2) | 1u;
        else if (op == OP_FALSE) st <<= 2;
        else if (op == OP_ERR) st = (st << 2) | 2u;
        else if (op == OP_AND3 || op == OP_OR3) {     // a = the slot below, b = the top slot
            const uint32_t vb = st & 1u, eb = (st >> 1) & 1u;
            st >>= 2;
            const uint32_t va = st & 1u, ea = (st >> 1) & 1u;
            const uint32_t v = op == OP_AND3 ? (va & vb) : (va | (vb & (ea ^ 1u)));
            const uint32_t e = op == OP_AND3 ? (ea | (va & eb)) : (ea | ((va ^ 1u) & eb));
            st = (st & ~3u) | v | (e << 1);
        } else if (op == OP_NOT3) st = (st & ~1u) | (((st | (st >> 1)) & 1u) ^ 1u);
        w = next;
    }
    err = err || (st & 2u) != 0u;
    return (st & 1u) != 0u;
}

}  // namespace bmc
