// tests/emul/cuda_shim/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
// A tiny stand-in for the CUDA execution model so that a .cu file of the product can be compiled with g++ and
// its kernels EXECUTED on the host by the CPU test-suite (there is no GPU in the dev container): every block
// runs as blockDim.x OS threads, __syncthreads is a block barrier, the warp intrinsics exchange values through
// a per-warp slot array guarded by a per-warp barrier.  Blocks run one after another.  Only what
// smash_paper_b200/csrc/ingest.cu uses is provided.  Never part of the product build (nvcc finds the real header).
#pragma once
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

typedef void *cudaStream_t;
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct ShimIdx { unsigned x = 0, y = 0, z = 0; };
inline thread_local ShimIdx threadIdx, blockIdx;
inline ShimIdx blockDim, gridDim;

struct ShimBlock {
  std::unique_ptr<std::barrier<>> block_barrier;
  std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
  std::vector<uint64_t> slots;          // 32 per warp
};
inline ShimBlock *g_shim = nullptr;

inline void __syncthreads() { g_shim->block_barrier->arrive_and_wait(); }
inline void __threadfence_system() {}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (sh & 31)); }

template <class T> inline T shim_shfl_up(T v, int d) {
  static_assert(sizeof(T) <= 8, "shim shuffles move up to 64 bits");
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  g_shim->slots[32 * warp + lane] = raw;
  g_shim->warp_barrier[warp]->arrive_and_wait();
  uint64_t got = (int)lane >= d ? g_shim->slots[32 * warp + lane - d] : raw;
  g_shim->warp_barrier[warp]->arrive_and_wait();
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
inline unsigned __shfl_up_sync(unsigned, unsigned v, int d) { return shim_shfl_up(v, d); }
inline unsigned long long __shfl_up_sync(unsigned, unsigned long long v, int d) { return shim_shfl_up(v, d); }
inline unsigned __ballot_sync(unsigned, bool pred) {
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  g_shim->slots[32 * warp + lane] = pred ? 1 : 0;
  g_shim->warp_barrier[warp]->arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (unsigned)(g_shim->slots[32 * warp + l] & 1) << l;
  g_shim->warp_barrier[warp]->arrive_and_wait();
  return m;
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}

// kernel<<<grid, block, 0, st>>>(args...)
template <class... KArgs, class... Args>
inline void shim_launch(void (*k)(KArgs...), unsigned grid, unsigned block, Args... args) {
  gridDim.x = grid; blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b) {
    ShimBlock blk;
    blk.block_barrier.reset(new std::barrier<>(block));
    const unsigned n_warps = (block + 31) / 32;
    for (unsigned w = 0; w < n_warps; ++w) {
      const unsigned lanes = w + 1 < n_warps ? 32 : block - 32 * w;
      blk.warp_barrier.emplace_back(new std::barrier<>(lanes));
    }
    blk.slots.assign(32 * n_warps, 0);
    g_shim = &blk;
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([=]() { threadIdx.x = t; blockIdx.x = b; k(args...); });
    for (auto &x : th) x.join();
    g_shim = nullptr;
  }
}
