// tests/emul/cuda_shim/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
// A small stand-in for the CUDA execution model and the slice of the runtime API this repository uses, so that the
// product's .cu files can be compiled with g++ and their kernels EXECUTED on the host by the CPU test-suite (there is
// no GPU in the dev container).  A block runs as blockDim.x OS threads (blocks one after another), __syncthreads is
// a block barrier, the *_sync warp intrinsics are rendezvous points keyed by (warp, mask) that exchange values
// through per-lane slots, "device memory" is host memory, streams are synchronous.  It finds indexing, scan,
// shuffle and synchronisation bugs; it says nothing about performance.  Never part of the product build (nvcc
// finds the real header) and never loaded by the product.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define SMASH_CUDA_SHIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

// ---- vector types ------------------------------------------------------------------------------------------
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct alignas(16) int4 { int x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct ShimIdx { unsigned x = 0, y = 0, z = 0; };
inline thread_local ShimIdx threadIdx, blockIdx;
inline ShimIdx blockDim, gridDim;
constexpr int warpSize = 32;

// ---- per-launch execution state ------------------------------------------------------------------------------
struct ShimWarp {
  std::mutex m;
  std::condition_variable cv;
  unsigned exited = 0;                                   // lanes whose thread has returned from the kernel (this block)
  struct Point { int arrived = 0; unsigned long gen = 0; uint64_t in[32]; uint64_t out[32]; unsigned pred_in = 0, pred_out = 0; };
  std::map<unsigned, Point> points;                      // one rendezvous per participation mask
};
struct ShimLaunch {
  std::vector<std::unique_ptr<std::barrier<>>> sync;     // __syncthreads of block b
  std::vector<ShimWarp> warps;
  alignas(16) uint8_t *dyn_smem = nullptr;
};
inline ShimLaunch *g_shim = nullptr;
inline uint8_t *shim_dynamic_smem() { return g_shim->dyn_smem; }

inline void __syncthreads() { g_shim->sync[blockIdx.x]->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// All lanes named by `mask` (that have not exited) deposit (value, pred) and leave with everybody's deposits.
inline void shim_rendezvous(unsigned mask, uint64_t value, bool pred, uint64_t out[32], unsigned *preds) {
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ShimWarp &w = g_shim->warps[warp];
  const unsigned block_lanes = blockDim.x - 32 * warp >= 32 ? 0xffffffffu : ((1u << (blockDim.x - 32 * warp)) - 1u);
  mask &= block_lanes;
  std::unique_lock<std::mutex> lk(w.m);
  ShimWarp::Point &p = w.points[mask];
  const unsigned long my_gen = p.gen;
  p.in[lane] = value;
  if (pred) p.pred_in |= 1u << lane;
  ++p.arrived;
  auto complete = [&]() { return p.arrived >= __builtin_popcount(mask & ~w.exited); };
  if (complete()) {
    memcpy(p.out, p.in, sizeof p.out); p.pred_out = p.pred_in; p.pred_in = 0; p.arrived = 0; ++p.gen;
    w.cv.notify_all();
  } else {
    w.cv.wait(lk, [&]() {
      if (p.gen != my_gen) return true;
      if (complete()) {                                  // a lane of the mask exited meanwhile: the last waiter closes the point
        memcpy(p.out, p.in, sizeof p.out); p.pred_out = p.pred_in; p.pred_in = 0; p.arrived = 0; ++p.gen;
        w.cv.notify_all();
        return true;
      }
      return false;
    });
  }
  if (out) memcpy(out, p.out, sizeof p.out);
  if (preds) *preds = p.pred_out & mask;
}
inline void shim_lane_exit() {
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ShimWarp &w = g_shim->warps[warp];
  { std::lock_guard<std::mutex> lk(w.m); w.exited |= 1u << lane; }
  w.cv.notify_all();
}

inline void __syncwarp(unsigned mask = 0xffffffffu) { shim_rendezvous(mask, 0, false, nullptr, nullptr); }
template <class T> inline T shim_pick(unsigned mask, T v, int src_lane, bool valid) {
  static_assert(sizeof(T) <= 8, "shim shuffles move up to 64 bits");
  uint64_t raw = 0, out[32];
  memcpy(&raw, &v, sizeof(T));
  shim_rendezvous(mask, raw, false, out, nullptr);
  if (!valid) return v;
  T r;
  memcpy(&r, &out[src_lane & 31], sizeof(T));
  return r;
}
#define SHIM_SHFL(T)                                                                                                   \
  inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {                                                     \
    const int lane = threadIdx.x & 31, base = lane & ~(width - 1);                                                     \
    return shim_pick<T>(m, v, base + (src & (width - 1)), true);                                                       \
  }                                                                                                                    \
  inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {                                               \
    const int lane = threadIdx.x & 31, base = lane & ~(width - 1);                                                     \
    return shim_pick<T>(m, v, lane - (int)d, lane - (int)d >= base);                                                   \
  }                                                                                                                    \
  inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {                                             \
    const int lane = threadIdx.x & 31, base = lane & ~(width - 1);                                                     \
    return shim_pick<T>(m, v, lane + (int)d, lane + (int)d < base + width);                                            \
  }                                                                                                                    \
  inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {                                                   \
    const int lane = threadIdx.x & 31;                                                                                 \
    return shim_pick<T>(m, v, lane ^ x, (lane ^ x) < ((lane & ~(width - 1)) + width));                                 \
  }
SHIM_SHFL(int)
SHIM_SHFL(unsigned)
SHIM_SHFL(long)
SHIM_SHFL(unsigned long)
SHIM_SHFL(long long)
SHIM_SHFL(unsigned long long)
SHIM_SHFL(float)
SHIM_SHFL(double)
#undef SHIM_SHFL
inline unsigned __ballot_sync(unsigned mask, int pred) { unsigned p; shim_rendezvous(mask, 0, pred != 0, nullptr, &p); return p; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
  const unsigned warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(mask, pred);
  unsigned live = mask;
  { std::lock_guard<std::mutex> lk(g_shim->warps[warp].m); live &= ~g_shim->warps[warp].exited; }
  return (b & live) == live;
}
template <class T> inline T shim_reduce_add(unsigned mask, T v) {
  uint64_t raw = 0, out[32];
  memcpy(&raw, &v, sizeof(T));
  shim_rendezvous(mask, raw, true, out, &mask);          // pred = "I took part": the sum runs over the lanes that arrived
  T s = 0;
  for (int l = 0; l < 32; ++l) if (mask & (1u << l)) { T x; memcpy(&x, &out[l], sizeof(T)); s += x; }
  return s;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return shim_reduce_add<unsigned>(mask, v); }
inline int __reduce_add_sync(unsigned mask, int v) { return shim_reduce_add<int>(mask, v); }
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  uint64_t out[32]; unsigned took;
  shim_rendezvous(mask, v, true, out, &took);
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) if (took & (1u << l)) m = out[l] > m ? (unsigned)out[l] : m;
  return m;
}

// ---- scalar intrinsics --------------------------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)(((((uint64_t)hi) << 32) | lo) >> (sh & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { return (unsigned)((((((uint64_t)hi) << 32) | lo) << (sh & 31)) >> 32); }
inline unsigned __vsadu4(unsigned a, unsigned b) {               // sum of absolute differences of the four unsigned bytes
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) { const int x = (int)((a >> (8 * i)) & 0xffu), y = (int)((b >> (8 * i)) & 0xffu); r += (unsigned)(x > y ? x - y : y - x); }
  return r;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  const uint64_t v = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }
template <class T> inline void __stcg(T *p, T v) { *p = v; }

// ---- atomics (global and shared memory alike) ------------------------------------------------------------------------
#define SHIM_ATOMICS(T)                                                                                                \
  inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }                                   \
  inline T atomicSub(T *p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }                                   \
  inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }                                     \
  inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }                                   \
  inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }                                 \
  inline T atomicCAS(T *p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return cmp; } \
  inline T atomicMin(T *p, T v) {                                                                                        \
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);                                                                        \
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}            \
    return old;                                                                                                          \
  }                                                                                                                      \
  inline T atomicMax(T *p, T v) {                                                                                        \
    T old = __atomic_load_n(p, __ATOMIC_RELAXED);                                                                        \
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}            \
    return old;                                                                                                          \
  }
SHIM_ATOMICS(int)
SHIM_ATOMICS(unsigned)
SHIM_ATOMICS(long long)
SHIM_ATOMICS(unsigned long)
SHIM_ATOMICS(unsigned long long)
#undef SHIM_ATOMICS

// ---- runtime API: device memory is host memory, streams are synchronous ---------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
typedef void *cudaStream_t;
struct ShimEvent { int dummy; };
typedef ShimEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaHostAllocDefault = 0, cudaHostAllocMapped = 2, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrComputeCapabilityMajor = 75 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };

inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255 + 64) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
template <class T> inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned f) { return cudaHostAlloc((void **)p, n, f); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(d, v, n); }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new ShimEvent(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.001f; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "shim error"; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) {
  *v = a == cudaDevAttrComputeCapabilityMajor ? 10 : 1;   // ONE "SM": grids sized from the SM count stay small
  return cudaSuccess;
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
  a->type = cudaMemoryTypeHost; a->device = 0; a->devicePointer = (void *)p; a->hostPointer = (void *)p;
  return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
enum cudaLimit { cudaLimitMaxL2FetchGranularity = 5 };
inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }

// ---- kernel<<<grid, block, smem, stream>>>(args...)  ==  shim_bind(kernel, grid, block, smem, stream)(args...) --------------
template <class... KArgs> struct ShimBound {
  void (*k)(KArgs...); unsigned grid, block; size_t smem;
  template <class... Args> void operator()(Args... args) const {
    if (!grid || !block) return;
    ShimLaunch L;
    gridDim.x = grid; blockDim.x = block;
    const unsigned n_warps = (block + 31) / 32;
    L.warps = std::vector<ShimWarp>(n_warps);
    for (unsigned b = 0; b < grid; ++b) L.sync.emplace_back(new std::barrier<>(block));
    std::vector<uint8_t> dyn(smem + 64);
    L.dyn_smem = (uint8_t *)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    std::barrier<> block_edge(block);                      // static __shared__ storage is reused: blocks run one after another
    g_shim = &L;
    void (*kernel)(KArgs...) = k;
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([=, &L, &block_edge]() {
        threadIdx.x = t;
        for (unsigned b = 0; b < grid; ++b) {
          blockIdx.x = b;
          kernel(args...);
          L.sync[b]->arrive_and_drop();                    // a returned thread no longer takes part in __syncthreads
          shim_lane_exit();
          block_edge.arrive_and_wait();
          if ((t & 31) == 0) { ShimWarp &w = L.warps[t >> 5]; std::lock_guard<std::mutex> lk(w.m); w.exited = 0; w.points.clear(); }
          block_edge.arrive_and_wait();
        }
      });
    for (auto &x : th) x.join();
    g_shim = nullptr;
  }
};
template <class... KArgs>
inline ShimBound<KArgs...> shim_bind(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = nullptr) {
  return ShimBound<KArgs...>{k, grid.x, block.x, smem};
}
// the form ingest.cu's ing_launch uses
template <class... KArgs, class... Args>
inline void shim_launch(void (*k)(KArgs...), unsigned grid, unsigned block, Args... args) { shim_bind(k, grid, block, 0, nullptr)(args...); }
