// sabuild.cu -- index construction on the GPU: SA, ISA, LCP(+overflow table) of the text, i.e.
// the arrays the reference builds with qsufsort + Kasai (longSA.cpp:137-176, qsufsort.cpp:266-344,
// computeLCP longSA.cpp:224-237).  All three are canonical functions of the text, so the files
// written from these arrays are byte-identical to the reference's.
//
// Method: MSD bucket sort.  Suffixes are partitioned by their first two bytes into chunks that fit
// the work buffers; inside a chunk they are radix-sorted (CUB -- index construction is not the hot
// path) on a 64-bit key of the next C characters (C = 64 / bits-per-symbol after alphabet
// compaction), then only the still-tied groups are re-keyed C characters deeper and re-sorted
// until every group is a singleton.  LCP comes from a direct comparison of neighbouring suffixes.
#include <cub/cub.cuh>
#include <stdio.h>

#include <vector>

#include "kernels.cuh"
#include "sabuild.cuh"

namespace smash {

#define SCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, 256, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return -3; } } while (0)

struct KeySpec { int bits, chars; uint8_t rank[256]; };
static inline int gridn(uint64_t n) { uint64_t g = (n + 255) / 256; return (int)(g < 148 * 16 ? (g ? g : 1) : 148 * 16); }

// pair id of the first two symbols of suffix i: rank pairs (32 x 32) when the alphabet is small,
// raw byte pairs otherwise.  Either way pair order == lexicographic order of the two symbols.
__device__ __forceinline__ unsigned pair_id(const uint8_t *__restrict__ T, uint64_t N, uint64_t i, const KeySpec &ks, bool small) {
  const uint8_t a = T[i], b = i + 1 < N ? T[i + 1] : 0;
  return small ? ((unsigned)ks.rank[a] << 5) | ks.rank[b] : ((unsigned)a << 8) | b;
}
__global__ void k_hist2(const uint8_t *__restrict__ T, uint64_t N, KeySpec ks, int small, unsigned long long *hist) {
  __shared__ unsigned int sh[1024];
  if (small) { for (int k = threadIdx.x; k < 1024; k += blockDim.x) sh[k] = 0; __syncthreads(); }
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned b = pair_id(T, N, i, ks, small);
    if (small) atomicAdd(&sh[b], 1u); else atomicAdd(&hist[b], 1ull);
  }
  if (small) { __syncthreads(); for (int k = threadIdx.x; k < 1024; k += blockDim.x) if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]); }
}
__global__ void k_select(const uint8_t *__restrict__ T, uint64_t N, KeySpec ks, int small, unsigned lo, unsigned hi,
                         uint64_t *__restrict__ vals, unsigned long long *counter) {
  for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < N; i0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = i0 + threadIdx.x;
    bool take = false;
    if (i < N) { const unsigned b = pair_id(T, N, i, ks, small); take = b >= lo && b < hi; }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m) {
      const int lane = threadIdx.x & 31;
      unsigned long long base = 0;
      if (lane == __ffs((int)m) - 1) base = atomicAdd(counter, (unsigned long long)__popc(m));
      base = __shfl_sync(0xffffffffu, base, __ffs((int)m) - 1);
      if (take) vals[base + __popc(m & ((1u << lane) - 1))] = i;
    }
  }
}
__global__ void k_iota(uint64_t *v, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] = i;
}
__global__ void k_iota32(uint32_t *v, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}
// key of the C symbols at text[pos+depth ..], first symbol in the most significant bits
__global__ void k_keys(const uint8_t *__restrict__ T, uint64_t N, KeySpec ks, const uint64_t *__restrict__ pos, uint64_t n,
                       uint64_t depth, uint64_t *__restrict__ keys) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = pos[i] + depth;
    uint64_t key = 0;
    int done = 0;
    while (done < ks.chars) {
      uint64_t wv = p + done < N + TEXT_PAD - 8 ? text8(T, (int64_t)(p + done)) : 0;
      for (int j = 0; j < 8 && done < ks.chars; ++j, ++done) { key = (key << ks.bits) | ks.rank[wv & 0xff]; wv >>= 8; }
    }
    keys[i] = key;
  }
}
// after a sort by key: flag elements that are NOT alone in their key group, and the group start
__global__ void k_groups0(const uint64_t *__restrict__ keys, uint64_t n, uint8_t *__restrict__ tied, uint64_t *__restrict__ headidx) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const bool head = i == 0 || keys[i] != keys[i - 1];
    const bool tail = i + 1 == n || keys[i + 1] != keys[i];
    tied[i] = !(head && tail);
    headidx[i] = head ? i : 0;
  }
}
__global__ void k_gather64(const uint64_t *__restrict__ src, const uint32_t *__restrict__ perm, uint64_t n, uint64_t *__restrict__ dst) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[perm[i]];
}
// U sorted by (gid, key): run starts by gid
__global__ void k_runheads(const uint64_t *__restrict__ gid, uint64_t n, uint64_t *__restrict__ runstart) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    runstart[i] = (i == 0 || gid[i] != gid[i - 1]) ? i : 0;
}
// place the re-sorted group members, flag the ones still tied and compute their new group id
__global__ void k_place(const uint64_t *__restrict__ gid, const uint64_t *__restrict__ key, const uint64_t *__restrict__ pos,
                        const uint64_t *__restrict__ runstart, uint64_t n, uint64_t *__restrict__ cur,
                        uint8_t *__restrict__ tied, uint64_t *__restrict__ newhead) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t target = gid[i] + (i - runstart[i]);
    cur[target] = pos[i];
    const bool head = i == 0 || gid[i] != gid[i - 1] || key[i] != key[i - 1];
    const bool tail = i + 1 == n || gid[i + 1] != gid[i] || key[i + 1] != key[i];
    tied[i] = !(head && tail);
    newhead[i] = head ? target : 0;
  }
}
template <typename SaT>
__global__ void k_store_sa(const uint64_t *__restrict__ cur, uint64_t n, SaT *__restrict__ sa) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) sa[i] = (SaT)cur[i];
}
template <typename SaT>
__global__ void k_isa(const SaT *__restrict__ sa, uint64_t N, SaT *__restrict__ isa) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) isa[sa[i]] = (SaT)i;
}
// LCP[i] = lcp(suffix SA[i-1], suffix SA[i]) (computeLCP, longSA.cpp:224-237), by direct comparison
template <typename SaT>
__device__ __forceinline__ uint64_t lcp_direct(const uint8_t *__restrict__ T, uint64_t N, const SaT *__restrict__ sa, uint64_t i) {
  if (!i) return 0;
  const uint64_t a = sa[i - 1], b = sa[i];
  const uint64_t lim = N - (a > b ? a : b);
  uint64_t h = 0;
  while (h < lim) {
    const uint64_t d = text8(T, (int64_t)(a + h)) ^ text8(T, (int64_t)(b + h));
    if (d) { h += (uint64_t)(ctz64(d) >> 3); break; }
    h += 8;
  }
  return h < lim ? h : lim;
}
template <typename SaT>
__global__ void k_lcp(const uint8_t *__restrict__ T, uint64_t N, const SaT *__restrict__ sa, uint8_t *__restrict__ vec,
                      uint8_t *__restrict__ big) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t h = lcp_direct(T, N, sa, i);
    vec[i] = (uint8_t)(h < 255 ? h : 255);
    big[i] = h >= 255;
  }
}
template <typename SaT>
__global__ void k_lcpm_fill(const uint8_t *__restrict__ T, uint64_t N, const SaT *__restrict__ sa, const uint64_t *__restrict__ idx,
                            uint64_t n, LcpItem *__restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    out[i].idx = idx[i]; out[i].val = lcp_direct(T, N, sa, idx[i]);
  }
}


int build_index_device(const uint8_t *T, uint64_t N, int w, void *d_sa, void *d_isa, uint8_t *d_lcp,
                       LcpItem **d_lcpm, uint64_t *n_m, uint64_t chunk_cap, cudaStream_t st, char *err,
                       uint64_t *launches) {
  // ---- alphabet -> compact symbol ranks
  KeySpec ks; memset(&ks, 0, sizeof ks);
  uint32_t *d_alpha = nullptr; uint32_t alpha[8];
  SCU(cudaMalloc((void **)&d_alpha, 32));
  *launches += launch_alpha(T, N, d_alpha, st);
  SCU(cudaMemcpyAsync(alpha, d_alpha, 32, cudaMemcpyDeviceToHost, st));
  SCU(cudaStreamSynchronize(st));
  cudaFree(d_alpha);
  int sigma = 0;
  for (int c = 0; c < 256; ++c) if ((alpha[c >> 5] >> (c & 31)) & 1u) ks.rank[c] = (uint8_t)(++sigma);
  ks.bits = 1; while ((1 << ks.bits) < sigma + 1) ++ks.bits;
  ks.chars = 64 / ks.bits;
  const int small = sigma <= 31;
  const unsigned NB = small ? 1024u : 65536u;
  unsigned long long *d_hist = nullptr;
  SCU(cudaMalloc((void **)&d_hist, 65536 * 8));
  SCU(cudaMemsetAsync(d_hist, 0, 65536 * 8, st));
  k_hist2<<<gridn(N) < 148 * 4 ? gridn(N) : 148 * 4, 256, 0, st>>>(T, N, ks, small, d_hist);
  std::vector<unsigned long long> hist(65536);
  SCU(cudaMemcpyAsync(hist.data(), d_hist, 65536 * 8, cudaMemcpyDeviceToHost, st));
  SCU(cudaStreamSynchronize(st));
  cudaFree(d_hist);
  *launches += 1;
  // ---- chunk plan over the 65536 two-byte buckets (lexicographic order == bucket order)
  if (chunk_cap == 0) chunk_cap = 1ull << 30;
  uint64_t maxm = 0;
  std::vector<unsigned> cuts; cuts.push_back(0);
  { uint64_t acc = 0;
    for (unsigned b = 0; b < NB; ++b) {
      if (acc && acc + hist[b] > chunk_cap) { cuts.push_back(b); if (acc > maxm) maxm = acc; acc = 0; }
      acc += hist[b];
    }
    cuts.push_back(NB); if (acc > maxm) maxm = acc; }
  const uint64_t M = maxm;
  uint64_t *keysA = nullptr, *keysB = nullptr, *valsA = nullptr, *valsB = nullptr, *gidA = nullptr, *gidB = nullptr, *aux = nullptr, *aux2 = nullptr;
  uint32_t *permA = nullptr, *permB = nullptr; uint8_t *tied = nullptr; void *tmp = nullptr; unsigned long long *d_cnt = nullptr;
  SCU(cudaMalloc((void **)&keysA, 8 * M)); SCU(cudaMalloc((void **)&keysB, 8 * M));
  SCU(cudaMalloc((void **)&valsA, 8 * M)); SCU(cudaMalloc((void **)&valsB, 8 * M));
  SCU(cudaMalloc((void **)&gidA, 8 * M)); SCU(cudaMalloc((void **)&gidB, 8 * M));
  SCU(cudaMalloc((void **)&aux, 8 * M)); SCU(cudaMalloc((void **)&aux2, 8 * M));
  SCU(cudaMalloc((void **)&permA, 4 * M)); SCU(cudaMalloc((void **)&permB, 4 * M));
  SCU(cudaMalloc((void **)&tied, M)); SCU(cudaMalloc((void **)&d_cnt, 16));
  size_t tb1 = 0, tb2 = 0, tb3 = 0, tb4 = 0, tb5 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb1, keysA, keysB, valsA, valsB, M, 0, 64, st);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, keysA, keysB, permA, permB, M, 0, 64, st);
  cub::DeviceScan::InclusiveScan(nullptr, tb3, aux, aux, cub::Max(), M, st);
  cub::DeviceSelect::Flagged(nullptr, tb4, valsA, tied, valsB, d_cnt, M, st);
  cub::DeviceSelect::Flagged(nullptr, tb5, cub::CountingInputIterator<uint64_t>(0), tied, valsB, d_cnt, N, st);
  size_t tbytes = tb1; if (tb2 > tbytes) tbytes = tb2; if (tb3 > tbytes) tbytes = tb3; if (tb4 > tbytes) tbytes = tb4; if (tb5 > tbytes) tbytes = tb5;
  SCU(cudaMalloc(&tmp, tbytes + 256));
  int mbits = 1; while ((1ull << mbits) < M + 1) ++mbits;

  uint64_t base = 0;
  for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
    uint64_t m = 0;
    for (unsigned b = cuts[ci]; b < cuts[ci + 1]; ++b) m += hist[b];
    if (!m) continue;
    if (cuts.size() == 2) { k_iota<<<gridn(N), 256, 0, st>>>(valsA, N); }
    else {
      SCU(cudaMemsetAsync(d_cnt, 0, 8, st));
      k_select<<<gridn(N), 256, 0, st>>>(T, N, ks, small, cuts[ci], cuts[ci + 1], valsA, d_cnt);
    }
    k_keys<<<gridn(m), 256, 0, st>>>(T, N, ks, valsA, m, 0, keysA);
    size_t tb = tbytes;
    SCU(cub::DeviceRadixSort::SortPairs(tmp, tb, keysA, keysB, valsA, valsB, m, 0, ks.bits * ks.chars, st));
    uint64_t *cur = valsB;                                    // the chunk's suffix array in progress
    k_groups0<<<gridn(m), 256, 0, st>>>(keysB, m, tied, aux);
    tb = tbytes; SCU(cub::DeviceScan::InclusiveScan(tmp, tb, aux, aux, cub::Max(), m, st));
    tb = tbytes; SCU(cub::DeviceSelect::Flagged(tmp, tb, aux, tied, gidA, d_cnt, m, st));      // group ids of tied elements
    tb = tbytes; SCU(cub::DeviceSelect::Flagged(tmp, tb, cur, tied, valsA, d_cnt, m, st));     // their positions
    unsigned long long u = 0;
    SCU(cudaMemcpyAsync(&u, d_cnt, 8, cudaMemcpyDeviceToHost, st)); SCU(cudaStreamSynchronize(st));
    *launches += 8;
    uint64_t depth = (uint64_t)ks.chars;
    // U = (gidA, valsA)[0..u): still-tied suffixes with the start index of their group in cur
    while (u) {
      k_keys<<<gridn(u), 256, 0, st>>>(T, N, ks, valsA, u, depth, keysA);
      k_iota32<<<gridn(u), 256, 0, st>>>(permA, u);
      tb = tbytes; SCU(cub::DeviceRadixSort::SortPairs(tmp, tb, keysA, keysB, permA, permB, u, 0, ks.bits * ks.chars, st));
      k_gather64<<<gridn(u), 256, 0, st>>>(gidA, permB, u, gidB);
      tb = tbytes; SCU(cub::DeviceRadixSort::SortPairs(tmp, tb, gidB, keysB /*sorted gid*/, permB, permA, u, 0, mbits, st));
      // now order = permA; sorted gid in keysB; gather key and pos
      k_gather64<<<gridn(u), 256, 0, st>>>(keysA, permA, u, gidB);       // gidB := key in final order
      k_gather64<<<gridn(u), 256, 0, st>>>(valsA, permA, u, aux2);       // aux2 := pos in final order
      k_runheads<<<gridn(u), 256, 0, st>>>(keysB, u, aux);
      tb = tbytes; SCU(cub::DeviceScan::InclusiveScan(tmp, tb, aux, aux, cub::Max(), u, st));
      k_place<<<gridn(u), 256, 0, st>>>(keysB, gidB, aux2, aux, u, cur, tied, keysA /*new head target*/);
      tb = tbytes; SCU(cub::DeviceScan::InclusiveScan(tmp, tb, keysA, keysA, cub::Max(), u, st));
      tb = tbytes; SCU(cub::DeviceSelect::Flagged(tmp, tb, keysA, tied, gidA, d_cnt, u, st));
      tb = tbytes; SCU(cub::DeviceSelect::Flagged(tmp, tb, aux2, tied, valsA, d_cnt, u, st));
      SCU(cudaMemcpyAsync(&u, d_cnt, 8, cudaMemcpyDeviceToHost, st)); SCU(cudaStreamSynchronize(st));
      *launches += 13;
      depth += (uint64_t)ks.chars;
    }
    if (w == 4) k_store_sa<uint32_t><<<gridn(m), 256, 0, st>>>(cur, m, (uint32_t *)d_sa + base);
    else k_store_sa<uint64_t><<<gridn(m), 256, 0, st>>>(cur, m, (uint64_t *)d_sa + base);
    *launches += 1;
    base += m;
  }
  SCU(cudaStreamSynchronize(st));
  void *fr[] = {keysA, keysB, valsA, valsB, gidA, gidB, aux, aux2, permA, permB};
  for (void *p : fr) cudaFree(p);
  // ---- ISA
  if (d_isa) {
    if (w == 4) k_isa<uint32_t><<<gridn(N), 256, 0, st>>>((const uint32_t *)d_sa, N, (uint32_t *)d_isa);
    else k_isa<uint64_t><<<gridn(N), 256, 0, st>>>((const uint64_t *)d_sa, N, (uint64_t *)d_isa);
    *launches += 1;
  }
  // ---- LCP + overflow table
  uint8_t *big = nullptr; uint64_t *bigidx = nullptr;
  SCU(cudaMalloc((void **)&big, N));
  if (w == 4) k_lcp<uint32_t><<<gridn(N), 256, 0, st>>>(T, N, (const uint32_t *)d_sa, d_lcp, big);
  else k_lcp<uint64_t><<<gridn(N), 256, 0, st>>>(T, N, (const uint64_t *)d_sa, d_lcp, big);
  SCU(cudaMalloc((void **)&bigidx, 8 * (N / 16 + 1024)));
  // ordered compaction of the indices with LCP >= 255 (sorted by idx, as vec_uchar::init leaves them);
  // done in pieces of 2^30 so the item count always fits CUB's 32-bit fast path
  unsigned long long nm = 0;
  const uint64_t big_cap = N / 16 + 1024;
  for (uint64_t o = 0; o < N; o += (1ull << 30)) {
    const uint64_t cnt = N - o < (1ull << 30) ? N - o : (1ull << 30);
    if (nm >= big_cap) break;
    // worst case this piece could add cnt entries; stop cleanly if that would overflow
    size_t tb = tbytes;
    unsigned long long piece = 0;
    uint64_t room = big_cap - nm;
    if (room < cnt) {
      // count first (cheap: flags only) to stay inside the buffer
      cudaError_t e0 = cub::DeviceReduce::Sum(tmp, tb, big + o, d_cnt, (int)cnt, st);
      if (e0 != cudaSuccess) { snprintf(err, 256, "reduce: %s", cudaGetErrorString(e0)); return -3; }
      SCU(cudaMemcpyAsync(&piece, d_cnt, 8, cudaMemcpyDeviceToHost, st)); SCU(cudaStreamSynchronize(st));
      if (piece > room) { nm = big_cap + 1; break; }
      tb = tbytes;
    }
    cudaError_t e = cub::DeviceSelect::Flagged(tmp, tb, cub::CountingInputIterator<uint64_t>(o), big + o, bigidx + nm, d_cnt, (int)cnt, st);
    if (e != cudaSuccess) { snprintf(err, 256, "select: %s", cudaGetErrorString(e)); return -3; }
    SCU(cudaMemcpyAsync(&piece, d_cnt, 8, cudaMemcpyDeviceToHost, st)); SCU(cudaStreamSynchronize(st));
    nm += piece;
  }
  if (nm > N / 16 + 1024) { snprintf(err, 256, "LCP overflow table too large (%llu entries)", nm); return -6; }
  SCU(cudaMalloc((void **)d_lcpm, sizeof(LcpItem) * (nm + 1)));
  if (nm) {
    if (w == 4) k_lcpm_fill<uint32_t><<<gridn(nm), 256, 0, st>>>(T, N, (const uint32_t *)d_sa, bigidx, nm, *d_lcpm);
    else k_lcpm_fill<uint64_t><<<gridn(nm), 256, 0, st>>>(T, N, (const uint64_t *)d_sa, bigidx, nm, *d_lcpm);
  }
  *launches += 3;
  SCU(cudaStreamSynchronize(st));
  SCU(cudaGetLastError());
  *n_m = nm;
  cudaFree(big); cudaFree(bigidx); cudaFree(tmp); cudaFree(tied); cudaFree(d_cnt);
  return 0;
}

}  // namespace smash
