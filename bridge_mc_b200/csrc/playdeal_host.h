// playdeal_host.h -- host-side entry to the trick estimator's kernels (playdeal_kernels.cu, built on playdeal.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bmc_pd {
constexpr int THREADS = 128, WARPS = THREADS / 32;      // one warp per deal
// sets the dynamic shared memory attribute; *blocks_per_sm = resident blocks per SM
cudaError_t configure(int *blocks_per_sm);
// idx == nullptr: deals 0..n-1; else the n deals idx[0..n).  d_probe != nullptr: no deal is played; d_probe[i] receives
// the deal's expected-length score (play_deal's probe mode) and d_out is not written
void launch(unsigned blocks, cudaStream_t stream, const void *d_records, const void *d_hands4, const unsigned int *d_idx, uint64_t n,
            const int8_t *d_trump, const uint8_t *d_declarer, int per_deal, uint8_t *d_arenas, uint32_t arena_bytes, void *d_out,
            unsigned long long *d_next, int timing, unsigned int *d_probe = nullptr);
// idx / count of the deals whose status is 1 (arena outgrown)
void launch_failed(cudaStream_t stream, const void *d_out, uint64_t n, unsigned int *d_idx, unsigned int *d_count);
}
