// sabuild.cuh -- GPU construction of the reference's index arrays (see sabuild.cu).
#pragma once
#include <cuda_runtime.h>
#include "core.cuh"
namespace smash {
// T: device text (padded as DevIndex::text). d_sa (N*w), d_isa (N*w or null), d_lcp (N) are
// caller-allocated; *d_lcpm is allocated here. chunk_cap = max suffixes sorted at once (0 = 2^30).
int build_index_device(const uint8_t *T, uint64_t N, int w, void *d_sa, void *d_isa, uint8_t *d_lcp,
                       LcpItem **d_lcpm, uint64_t *n_m, uint64_t chunk_cap, cudaStream_t st, char *err,
                       uint64_t *launches);
}
