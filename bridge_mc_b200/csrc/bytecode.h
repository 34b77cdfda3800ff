// bytecode.h -- feature indices and the rule bytecode, shared by the host compiler (rules.hpp) and the
// device interpreter (features.cuh).  Plain C++ with no library dependency so that it can also be fed to NVRTC.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

namespace bmc {

// ---- feature vector --------------------------------------------------------
enum Feature : int { F_C = 0, F_D, F_H, F_S, F_CP, F_DP, F_HP, F_SP, F_HCP, F_LOS, F_BAL, F_COUNT };

// ---- bytecode ----------------------------------------------------------------
// word: bits 0-3 op | 4-6 cmp | 8-11 feature a | 12-15 feature b | 16-23 const
// OP_ERR pushes the ERROR value -- a form whose head is not a function (the un-spliced rule list of
// bidding.cl:249): Lisp signals "illegal function call" if and only if evaluation REACHES it.  Programs
// that hold an OP_ERR use the three-valued OP_AND3 / OP_OR3 / OP_NOT3 (value bit + error bit per stack
// slot, Lisp's left-to-right short-circuit: and(a,b) = a is error -> error; a false -> false; else b),
// so an error inside a branch that `and` / `or` never evaluate does not surface; all other programs
// keep the two-valued ops.
enum Op : uint32_t { OP_END = 0, OP_CMPC, OP_CMPF, OP_AND, OP_OR, OP_NOT, OP_TRUE, OP_FALSE,
                     OP_ERR, OP_AND3, OP_OR3, OP_NOT3 };
// cmp field = 3-bit truth table over (a<b, a==b, a>b)
enum Cmp : uint32_t { CMP_LT = 1, CMP_LE = 3, CMP_GT = 4, CMP_GE = 6, CMP_EQ = 2, CMP_NE = 5 };

#ifndef __CUDACC_RTC__
inline uint32_t enc(Op op, Cmp c = CMP_LT, int a = 0, int b = 0, int k = 0)
{
    return (uint32_t)op | ((uint32_t)c << 4) | ((uint32_t)a << 8) | ((uint32_t)b << 12) | ((uint32_t)(k & 0xFF) << 16);
}
#endif

}  // namespace bmc
