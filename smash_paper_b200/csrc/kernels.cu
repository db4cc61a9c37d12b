// kernels.cu -- hand-written sm_100a kernels of the SMASH hot path and their launchers.
//
//   k_uniq_build / k_seed_build / k_alpha   index-derived structures, once per context
//   k_mam_search     K1: maximal almost-unique matches (longSA::MAM, longSA.cpp:503-546)
//   k_records        K3: resolve + merge + CIGAR items + XE + HI order (query.cpp:68-97, 231-320)
//   k_sizes + scan   K4a: exact SAM bytes per read and their exclusive prefix sum
//   k_emit           K4b: SAM text (print_matches, query.cpp:331-415) at the scanned offsets
//
// One warp owns one read in every per-read kernel; the read is staged once in shared memory.
// All kernels are grid-stride over reads with a grid sized from the SM count.
#include "kernels.cuh"
#if !defined(SMASH_CUDA_SHIM)
#include <cub/device/device_merge_sort.cuh>
#else
#include <algorithm>
#endif

namespace smash {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int PBUF = P_FRONT + MAXQ_FAST + P_BACK + 8;     // + word rounding of the word-wise staging
constexpr int SCR_CAP = 64;       // alignments per read the shared-memory scratch holds
constexpr int LANE_CAP = 4;
constexpr int LINE_BUF = 1024;

static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}
// grid caps (CTAs per SM) of the thread-per-item kernels; A/B knobs of the variant builds.  k_rec_build and k_sizes have
// very uneven items (0..10 matches per read, records of an unmapped read vs a five-fragment one): 64 CTAs per SM instead
// of 16 lets the block scheduler even them out (records 1.140 -> 1.115 ms, sizes+scan 0.290 -> 0.245 ms, r02w A/B);
// k_emit_text gains nothing from it (it only moves the tail append beside k_emit_copy instead).
#ifndef SMASH_RECB_CTAS
#define SMASH_RECB_CTAS 64
#endif
#ifndef SMASH_SIZES_CTAS
#define SMASH_SIZES_CTAS 64
#endif
// k_rec_xe is a persistent warp-per-read kernel: its grid must be a whole number of waves.  6 CTAs per SM are resident
// (40 registers x 256 threads); the former grid of 8 per SM ran 1.33 waves, the last one on a third of the machine
// (records stage 1.12 -> 0.925 ms with 6, 12 or 24 CTAs per SM, r02y A/B).
#ifndef SMASH_XE_CTAS
#define SMASH_XE_CTAS 6
#endif
#ifndef SMASH_TEXT_CTAS
#define SMASH_TEXT_CTAS 16
#endif
static int grid_for_warps(uint64_t n_items, int ctas_per_sm) {
  uint64_t need = (n_items + WARPS - 1) / WARPS;
  uint64_t cap = (uint64_t)sm_count() * (uint64_t)ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------ derived index structures

__global__ void k_alpha(const uint8_t *__restrict__ text, uint64_t N, uint32_t *alpha8) {
  __shared__ uint32_t bm[8];
  if (threadIdx.x < 8) bm[threadIdx.x] = 0;
  __syncthreads();
  uint32_t loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
    uint8_t c = text[i];
    loc[c >> 5] |= 1u << (c & 31);
  }
  for (int j = 0; j < 8; ++j) if (loc[j]) atomicOr(&bm[j], loc[j]);
  __syncthreads();
  if (threadIdx.x < 8 && bm[threadIdx.x]) atomicOr(&alpha8[threadIdx.x], bm[threadIdx.x]);
}

// U[SA[i]] = min(255, max(LCP[i], LCP[i+1]) + 1): shortest unique prefix length of the suffix.
__global__ void k_uniq_build(DevIndex ix, uint8_t *__restrict__ uniq) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ix.N; i += (uint64_t)gridDim.x * blockDim.x) {
    int a = ix.lcp[i];
    int b = (i + 1 < ix.N) ? ix.lcp[i + 1] : 0;
    int m = a > b ? a : b;
    uniq[sa_at(ix, i)] = (uint8_t)(m >= 254 ? 255 : m + 1);
  }
}

// f(i) = number of k-mers (as strings over a<c<g<t) that are <= suffix SA[i]; monotone in i.
__device__ __forceinline__ uint64_t kmers_le_suffix(const DevIndex &ix, uint64_t i, int k) {
  const uint64_t s = sa_at(ix, i);
  uint64_t code = 0;
  for (int j = 0; j < k; ++j) {
    const uint64_t p = s + (uint64_t)j;
    const uint8_t ch = p < ix.N ? ix.text[p] : 0;
    const int b = base_code(ch);
    if (b <= 3) { code = (code << 2) | (uint64_t)b; continue; }
    // non-acgt byte: how many of a,c,g,t sort below it (ASCII order)
    const uint64_t below = ch < 'a' ? 0 : ch < 'c' ? 1 : ch < 'g' ? 2 : ch < 't' ? 3 : 4;
    return ((code << 2) + below) << (2 * (k - j - 1));
  }
  return code + 1;
}
// S[x] = first SA index whose suffix is >= k-mer x  (= number of suffixes < x), x in [0, 4^k].
template <typename SeedT>
__global__ void k_seed_build(DevIndex ix, SeedT *__restrict__ seed, int k) {
  const uint64_t total = ix.N + 1;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t prev = i ? kmers_le_suffix(ix, i - 1, k) : 0;
    const uint64_t cur = i < ix.N ? kmers_le_suffix(ix, i, k) : ((1ull << (2 * k)) + 1);
    for (uint64_t x = prev; x < cur; ++x) seed[x] = (SeedT)i;
  }
}

// The 8-byte seed table rewritten IN PLACE into its blocked form (core.cuh): thread B turns S[16B .. 16B+15] into block B.
// The only word it reads outside its own block is S[16B+16], the first word of block B+1, whose low 40 bits are S[16B+16]
// before AND after that block's conversion (aligned 8-byte accesses do not tear), so no copy of the table is needed.
__global__ void k_seed_irr_count(const uint64_t *__restrict__ seed, uint64_t n_blocks, unsigned long long *n_irr) {
  for (uint64_t B = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; B < n_blocks; B += (uint64_t)gridDim.x * blockDim.x) {
    bool irr = false;
    for (int i = 0; i < 16; ++i) irr |= seed[16 * B + i + 1] - seed[16 * B + i] >= 255;
    if (irr) atomicAdd(n_irr, 1ull);
  }
}
__global__ void k_seed_block(uint64_t *seed, uint64_t n_blocks, const uint32_t *__restrict__ ext, uint64_t N,
                             uint64_t *__restrict__ irr, unsigned long long *n_irr) {
  for (uint64_t B = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; B < n_blocks; B += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t s[17];
    for (int i = 0; i < 16; ++i) s[i] = seed[16 * B + i];
    s[16] = *(volatile const uint64_t *)(seed + 16 * B + 16) & SEED_BASE_MASK;
    bool irregular = false;
    for (int i = 0; i < 16; ++i) irregular |= s[i + 1] - s[i] >= 255;
    uint64_t wd[16];
    wd[0] = s[0]; wd[1] = 0; wd[2] = 0;
    if (irregular) {
      const uint64_t row = (uint64_t)atomicAdd(n_irr, 1ull);
      for (int i = 0; i < 17; ++i) irr[row * 17 + i] = s[i];
      wd[0] |= 1ull << 63; wd[1] = row;
    } else {
      for (int i = 0; i < 8; ++i) { wd[1] |= (s[i + 1] - s[i]) << (8 * i); wd[2] |= (s[i + 9] - s[i + 8]) << (8 * i); }
    }
    uint32_t e[SEED_INLINE];
    for (int j = 0; j < SEED_INLINE; ++j) e[j] = s[0] + (uint64_t)j < N ? ext[s[0] + (uint64_t)j] : 0u;
    for (int j = 0; j < SEED_INLINE / 2; ++j) wd[3 + j] = (uint64_t)e[2 * j] | ((uint64_t)e[2 * j + 1] << 32);
    uint4 *out = reinterpret_cast<uint4 *>(seed + 16 * B);
    for (int j = 0; j < 8; ++j) out[j] = make_uint4((uint32_t)wd[2 * j], (uint32_t)(wd[2 * j] >> 32), (uint32_t)wd[2 * j + 1], (uint32_t)(wd[2 * j + 1] >> 32));
  }
}
int launch_seed_irr_count(const void *seed, uint64_t n_blocks, unsigned long long *n_irr, cudaStream_t st) {
  k_seed_irr_count<<<sm_count() * 8, 256, 0, st>>>((const uint64_t *)seed, n_blocks, n_irr);
  return 1;
}
int launch_seed_block(void *seed, uint64_t n_blocks, const uint32_t *ext, uint64_t N, uint64_t *irr, unsigned long long *n_irr, cudaStream_t st) {
  k_seed_block<<<sm_count() * 8, 256, 0, st>>>((uint64_t *)seed, n_blocks, ext, N, irr, n_irr);
  return 1;
}

__global__ void k_ext_build(DevIndex ix, int k, uint32_t *__restrict__ ext) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ix.N; i += (uint64_t)gridDim.x * blockDim.x)
    ext[i] = ext_entry(ix.text, ix.N, sa_at(ix, i), k);
}
int launch_ext_build(const DevIndex &ix, int k, uint32_t *ext, cudaStream_t st) {
  k_ext_build<<<sm_count() * 8, 256, 0, st>>>(ix, k, ext);
  return 1;
}
int launch_alpha(const uint8_t *text, uint64_t N, uint32_t *alpha8, cudaStream_t st) {
  cudaMemsetAsync(alpha8, 0, 32, st);
  k_alpha<<<sm_count() * 4, 256, 0, st>>>(text, N, alpha8);
  return 1;
}
int launch_uniq_build(const DevIndex &ix, uint8_t *uniq, cudaStream_t st) {
  k_uniq_build<<<sm_count() * 8, 256, 0, st>>>(ix, uniq);
  return 1;
}
int launch_seed_build(const DevIndex &ix, void *seed, int k, int seed_w, cudaStream_t st) {
  if (seed_w == 4) k_seed_build<uint32_t><<<sm_count() * 8, 256, 0, st>>>(ix, (uint32_t *)seed, k);
  else k_seed_build<uint64_t><<<sm_count() * 8, 256, 0, st>>>(ix, (uint64_t *)seed, k);
  return 1;
}

// ------------------------------------------------------------------ read staging

// Warp-cooperative: lower-case the read into shared memory (with the pads core.cuh relies on)
// and report whether it holds a non-acgt byte that also occurs in the text (=> exact path).
__device__ __forceinline__ bool stage_read(const DevIndex &ix, const uint8_t *__restrict__ seq, int q,
                                           int nucleotides_only, uint8_t *pbuf, int lane) {
  uint8_t *P = pbuf + P_FRONT;
  bool odd = false;
  for (int j = lane; j < q; j += 32) {
    const uint8_t c = query_char(seq[j], nucleotides_only);
    P[j] = c;
    if (base_code(c) > 3 && in_alpha(ix, c)) odd = true;
  }
  if (lane < P_FRONT) pbuf[lane] = 0xFE;
  if (lane < P_BACK) P[q + lane] = 0xFF;
  __syncwarp();
  return __any_sync(0xffffffffu, odd);
}

// ------------------------------------------------------------------ K1: MAM search
//
// One warp per read.  (1) lanes = anchors: SWAR k-mer code from the staged read, seed-table lookup,
// bucket size.  (2) every (anchor, suffix) candidate of the read becomes one task in a shared-memory
// list (warp prefix sum) so that (3) lanes = candidates: all lanes extend different candidates at
// once -- SA entry, text words and the U byte of different candidates are in flight together and no
// lane waits for a neighbour with a bigger bucket.  (4) matches are rank-sorted by query offset and
// written in order.
constexpr int TASK_CAP = 128;

struct SearchSmem {
  uint16_t lut[256];                           // byte -> lower-cased byte | 0x100 (not acgt) | 0x200 (not acgt but occurs in the text)
  uint8_t pbuf[WARPS][PBUF];
  uint32_t inv[WARPS][MAXQ_FAST / 32 + 12];    // non-acgt mask, one bit per byte of the staging buffer (bit 0 = pbuf[P_FRONT])
  Match stage[WARPS][STAGE_CAP];
  uint64_t task_sa[WARPS][TASK_CAP];           // SA index of the candidate
  uint16_t task_x[WARPS][TASK_CAP];            // its anchor position
  int nstage[WARPS];
};

__device__ __forceinline__ void stage_push(SearchSmem &sm, int warp, const Match &m) {
  const int slot = atomicAdd(&sm.nstage[warp], 1);
  if (slot < STAGE_CAP) sm.stage[warp][slot] = m;
}
// Ordered emission of a warp's staged matches: rank by query offset.  A query offset carries at most ONE reportable
// match (two unique maximal matches with the same start would make the shorter one a repeat), but it can be STAGED more
// than once: every candidate of a saturated repeat family (U == 255, diagonal >= 255) falls back to exact_start at
// the same start and finds the same match.  Copies are dropped here; returns the number of distinct matches written
// (those with rank < cap).  All lanes of the warp must call it.
__device__ __forceinline__ int emit_ranked(const Match *stage, int ns, Match *dst, int cap, int lane) {
  static_assert(STAGE_CAP <= 64, "two ballots cover the stage");
  bool f0 = false, f1 = false;
  if (lane < ns) {
    f0 = true;
    const uint32_t qp = stage[lane].qpos;
    for (int f = 0; f < lane; ++f) if (stage[f].qpos == qp) { f0 = false; break; }
  }
  if (lane + 32 < ns) {
    f1 = true;
    const uint32_t qp = stage[lane + 32].qpos;
    for (int f = 0; f < lane + 32; ++f) if (stage[f].qpos == qp) { f1 = false; break; }
  }
  const unsigned m0 = __ballot_sync(0xffffffffu, f0), m1 = __ballot_sync(0xffffffffu, f1);
  for (int e = lane; e < ns; e += 32) {
    if (!(((e < 32 ? m0 >> e : m1 >> (e - 32)) & 1u))) continue;
    const Match me = stage[e];
    int rank = 0;
    for (int f = 0; f < ns; ++f) rank += (stage[f].qpos < me.qpos && (((f < 32 ? m0 >> f : m1 >> (f - 32)) & 1u) != 0)) ? 1 : 0;
    if (rank < cap) dst[rank] = me;
  }
  return __popc(m0) + __popc(m1);
}
// every start of the read through the exact per-start search (reads the anchor path cannot take)
__device__ __noinline__ void exact_all(const DevIndex &ix, SearchSmem &sm, int warp, int lane, const uint8_t *P, int q, const SearchParams &sp) {
  for (int p = lane; p + (int)sp.L <= q; p += 32) {
    Match m;
    if (exact_start(ix, P, q, p, sp.L, &m)) stage_push(sm, warp, m);
  }
}
// lanes = tasks: 4+4 pre-filter on every queued candidate; the survivors are parked for k_mam_verify
// (w.surv) or, when the read's parking row is full / absent, extended right here
__device__ __forceinline__ void run_tasks(const DevIndex &ix, SearchSmem &sm, int warp, int lane, const uint8_t *P, int q,
                                          const SearchParams &sp, int ntask, const WorkDev &w, uint64_t read, int &nsurv) {
  for (int t0 = 0; t0 < ntask; t0 += 32) {
    const int t = t0 + lane;
    uint64_t i = 0; int x = 0;
    bool pass = false;
    if (t < ntask) {
      i = sm.task_sa[warp][t]; x = (int)sm.task_x[warp][t];
      pass = !ix.ext || ext_may_reach(ix.ext[i], read_ext_codes(P, x, sp.k), sp.k, sp.L);   // 4-byte pre-filter (conservative form)
    }
    if (w.surv) {
      const unsigned mask = __ballot_sync(0xffffffffu, pass);
      const int np = __popc(mask);
      if (nsurv + np <= SURV_CAP) {
        if (pass) w.surv[read * SURV_CAP + nsurv + __popc(mask & ((1u << lane) - 1u))] = ((uint64_t)x << 48) | (8ull << 40) | i;   // left extension not known here
        nsurv += np;
        continue;
      }
    }
    if (pass) {
      const uint64_t c = sa_at(ix, i);
      Match m; int pl = 0;
      const int r = candidate_check(ix, P, q, x, sp.s, sp.k, sp.L, c, &m, &pl);
      if (r > 0) stage_push(sm, warp, m);
      else if (r < 0 && exact_start(ix, P, q, pl, sp.L, &m)) stage_push(sm, warp, m);
    }
  }
  __syncwarp();
}

// Warp-cooperative staging for the search: lower-case the read into shared memory through the
// per-CTA lookup table, build the non-acgt bit mask (one ballot per 32 bases) in the same pass, and
// report whether a non-acgt byte that also occurs in the text was seen (=> exact path).
__device__ __forceinline__ bool stage_read_masked(SearchSmem &sm, int warp, const uint8_t *__restrict__ seq, int q, int lane,
                                                  uint8_t *__restrict__ lc) {
  uint8_t *pbuf = sm.pbuf[warp];
  uint8_t *P = pbuf + P_FRONT;
  bool odd = false;
  const int rounds = q / 32 + 2;                 // one spare mask word (kmer_invalid reads word+1)
  for (int it = 0; it < rounds; ++it) {
    const int j = it * 32 + lane;
    uint16_t e = 0;
    if (j < q) { e = sm.lut[seq[j]]; P[j] = (uint8_t)e; odd |= (e & 0x200) != 0; if (lc) lc[j] = (uint8_t)e; }
    const unsigned bad = __ballot_sync(0xffffffffu, (e & 0x100) != 0);
    if (lane == 0) sm.inv[warp][it] = bad;
  }
  if (lane < P_FRONT) { pbuf[lane] = 0xFE; if (lc) lc[lane - P_FRONT] = 0xFE; }
  if (lane < P_BACK) { P[q + lane] = 0xFF; if (lc) lc[q + lane] = 0xFF; }
  __syncwarp();
  return __any_sync(0xffffffffu, odd);
}


// Word-wise staging (K1a): the read is fetched as aligned 32-bit words, ALL of a 256-byte pass in
// flight before the first one is used (one global latency per pass instead of one per 32 bytes),
// lower-cased through the per-CTA table four bytes at a time and stored as words.  The staged read
// starts at pbuf + P_FRONT + mis (mis = global misalignment of the read), so word boundaries of the
// global blob, the shared buffer and the lower-cased HBM copy (lc) coincide.  The non-acgt mask is
// kept in BUFFER coordinates (bit j + mis for base j).  Pads: 0xFE before, 0xFF after (core.cuh).
__device__ __forceinline__ uint32_t lut4(const uint16_t *lut, uint32_t v, int j0, int q, unsigned &nib, unsigned &oddbits) {
  const uint32_t e0 = lut[v & 0xffu], e1 = lut[(v >> 8) & 0xffu], e2 = lut[(v >> 16) & 0xffu], e3 = lut[v >> 24];
  if (j0 >= 0 && j0 + 3 < q) {                    // whole word inside the read
    nib = ((e0 >> 8) & 1u) | ((e1 >> 7) & 2u) | ((e2 >> 6) & 4u) | ((e3 >> 5) & 8u);
    oddbits |= (e0 | e1 | e2 | e3) & 0x200u;
    return (e0 & 0xffu) | ((e1 & 0xffu) << 8) | ((e2 & 0xffu) << 16) | (e3 << 24);
  }
  uint32_t out = 0; nib = 0;
  const uint32_t e[4] = {e0, e1, e2, e3};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int j = j0 + t;
    uint32_t byte;
    if (j < 0) byte = 0xFEu;
    else if (j >= q) byte = 0xFFu;
    else { byte = e[t] & 0xffu; nib |= ((e[t] >> 8) & 1u) << t; oddbits |= e[t] & 0x200u; }
    out |= byte << (8 * t);
  }
  return out;
}
__device__ __forceinline__ bool stage_read_words(SearchSmem &sm, int warp, const uint8_t *__restrict__ seq_blob, int64_t so, int q,
                                                 int lane, uint8_t *__restrict__ lc) {
  const int mis = (int)(so & 3);
  uint8_t *pbuf = sm.pbuf[warp];
  uint32_t *pw = reinterpret_cast<uint32_t *>(pbuf + P_FRONT);                 // word 0 holds bases -mis .. 3-mis
  const uint32_t *gw = reinterpret_cast<const uint32_t *>(seq_blob + (so - mis));
  uint32_t *lw = lc ? reinterpret_cast<uint32_t *>(lc - mis) : nullptr;
  const int nwords = (mis + q + 3) >> 2;
  unsigned oddbits = 0;
  int w0 = 0;
  for (; w0 < nwords; w0 += 64) {
    const int i0 = w0 + lane, i1 = w0 + 32 + lane;
    const uint32_t v0 = i0 < nwords ? __ldg(gw + i0) : 0u;
    const uint32_t v1 = i1 < nwords ? __ldg(gw + i1) : 0u;
    unsigned n0 = 0, n1 = 0;
    if (i0 < nwords) { const uint32_t o = lut4(sm.lut, v0, 4 * i0 - mis, q, n0, oddbits); pw[i0] = o; if (lw) lw[i0] = o; }
    if (i1 < nwords) { const uint32_t o = lut4(sm.lut, v1, 4 * i1 - mis, q, n1, oddbits); pw[i1] = o; if (lw) lw[i1] = o; }
    // 8 lanes x 4 flag bits -> one mask word per 32 buffer bytes
    unsigned m0 = n0 << (4 * (lane & 7)), m1 = n1 << (4 * (lane & 7));
    m0 |= __shfl_xor_sync(0xffffffffu, m0, 1); m1 |= __shfl_xor_sync(0xffffffffu, m1, 1);
    m0 |= __shfl_xor_sync(0xffffffffu, m0, 2); m1 |= __shfl_xor_sync(0xffffffffu, m1, 2);
    m0 |= __shfl_xor_sync(0xffffffffu, m0, 4); m1 |= __shfl_xor_sync(0xffffffffu, m1, 4);
    if ((lane & 7) == 0) { sm.inv[warp][i0 >> 3] = m0; sm.inv[warp][i1 >> 3] = m1; }
  }
  if (lane == 0) { sm.inv[warp][w0 >> 3] = 0; sm.inv[warp][(w0 >> 3) + 1] = 0; }   // spare words (kmer_invalid reads word+1)
  // pads outside the staged words
  if (lane < P_FRONT) { pbuf[lane] = 0xFE; if (lc && lane - P_FRONT < -mis) lc[lane - P_FRONT] = 0xFE; }
  {
    const int j = 4 * nwords - mis + lane;         // first byte after the last staged word, in read coordinates
    if (j < q + P_BACK) { pbuf[P_FRONT + mis + j] = 0xFF; if (lc) lc[j] = 0xFF; }
  }
  __syncwarp();
  return __any_sync(0xffffffffu, oddbits != 0);
}

#ifndef SMASH_MINBLK
#define SMASH_MINBLK 4
#endif
#ifndef SMASH_LUT
#define SMASH_LUT 2
#endif
__global__ void __launch_bounds__(THREADS, SMASH_MINBLK)
k_mam_search(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  __shared__ __align__(16) SearchSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const uint8_t c = query_char((uint8_t)threadIdx.x, sp.nucleotides_only);
    const bool not_acgt = base_code(c) > 3;
    sm.lut[threadIdx.x] = (uint16_t)(c | (not_acgt ? 0x100 : 0) | (not_acgt && in_alpha(ix, c) ? 0x200 : 0));
  }
  __syncthreads();
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t read = (uint64_t)blockIdx.x * WARPS + warp; read < b.n_reads; read += warps_total) {
    if (w.slow && !w.slow[read]) continue;         // k_mam_seed finished this read
    const int64_t so = b.seq_off[read];
    const int q = (int)(b.seq_off[read + 1] - so);
    if (lane == 0) sm.nstage[warp] = 0;
    if (q > MAXQ_FAST) {                           // k_mam_search_long takes these (the host learns the length from the flag)
      if (lane == 0) { atomicMax(&w.flags[FLAG_LONGQ], (uint32_t)q); if (q > w.long_q) w.match_cnt[read] = 0; if (w.surv) w.surv_cnt[read] = 0; }
      continue;
    }
    int nsurv = 0;
    {
#if SMASH_LUT == 2
    const bool odd = stage_read_words(sm, warp, b.seq, so, q, lane, w.surv ? w.lc + so + 32 * read + 16 : nullptr);
    const int mis = (int)(so & 3);
    const uint8_t *P = sm.pbuf[warp] + P_FRONT + mis;
#elif SMASH_LUT
    const bool odd = stage_read_masked(sm, warp, b.seq + so, q, lane, w.surv ? w.lc + so + 32 * read + 16 : nullptr);
    const int mis = 0;
    const uint8_t *P = sm.pbuf[warp] + P_FRONT;
#else
    const int mis = 0;
    const bool odd = stage_read(ix, b.seq + so, q, sp.nucleotides_only, sm.pbuf[warp], lane);
    const uint8_t *P = sm.pbuf[warp] + P_FRONT;
    for (int c0 = 0; c0 <= q / 32 + 1; ++c0) {
      const int j = c0 * 32 + lane;
      const unsigned bad = __ballot_sync(0xffffffffu, j < q && base_code(P[j]) > 3);
      if (lane == 0) sm.inv[warp][c0] = bad;
    }
    __syncwarp();
#endif
    const int L = (int)sp.L;
    if (q >= L) {
      if (!odd && sp.fast_ok) {
        // (bytes flagged in the non-acgt mask do not occur in the text here, so no match can contain them)
        const int s = sp.s, k = sp.k;
        const int n_anchor = (q - L + s - 1) / s + 1;        // anchors x = a*s cover starts 0..q-L
        int ntask = 0;
        for (int a0 = 0; a0 < n_anchor; a0 += 32) {
          const int a = a0 + lane;
          uint64_t lo = 0, hi = 0;
          bool big = false;
          if (a < n_anchor && !kmer_invalid(sm.inv[warp], a * s + mis, k)) {
            anchor_bucket(ix, P, a * s, k, &lo, &hi);
            if (hi - lo > (uint64_t)BIG_BUCKET) { big = true; hi = lo; }
          }
          unsigned slow = __ballot_sync(0xffffffffu, big);
          while (slow) {                                     // huge bucket: exact per-start search for that window
            const int la = __ffs((int)slow) - 1; slow &= slow - 1;
            const int x = (a0 + la) * s;
            const int p_lo = x - s + 1 > 0 ? x - s + 1 : 0;
            for (int p = p_lo + lane; p <= x; p += 32) {
              Match m;
              if (exact_start(ix, P, q, p, sp.L, &m)) stage_push(sm, warp, m);
            }
          }
          // queue this round's candidates; flush the list whenever the round would not fit
          int cnt = (int)(hi - lo);
          int inc = cnt;
          for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
          const int total = __shfl_sync(0xffffffffu, inc, 31);
          if (ntask + total > TASK_CAP) { run_tasks(ix, sm, warp, lane, P, q, sp, ntask, w, read, nsurv); ntask = 0; }
          if (total > TASK_CAP) {                            // cannot happen: 32 * BIG_BUCKET > TASK_CAP only if many full buckets
            for (int l2 = 0; l2 < 32; ++l2) {                // degrade gracefully: one lane's bucket at a time
              const int c2 = __shfl_sync(0xffffffffu, cnt, l2);
              const uint64_t lo2 = __shfl_sync(0xffffffffu, lo, l2);
              for (int i = lane; i < c2; i += 32) { sm.task_sa[warp][i] = lo2 + (uint64_t)i; sm.task_x[warp][i] = (uint16_t)((a0 + l2) * s); }
              __syncwarp();
              run_tasks(ix, sm, warp, lane, P, q, sp, c2, w, read, nsurv);
            }
          } else {
            const int base = ntask + inc - cnt;
            for (int i = 0; i < cnt; ++i) { sm.task_sa[warp][base + i] = lo + (uint64_t)i; sm.task_x[warp][base + i] = (uint16_t)(a * s); }
            ntask += total;
            __syncwarp();
          }
        }
        run_tasks(ix, sm, warp, lane, P, q, sp, ntask, w, read, nsurv);
      } else {
        exact_all(ix, sm, warp, lane, P, q, sp);
      }
    }
    }   // q <= MAXQ_FAST
    __syncwarp();
    const int n_raw = sm.nstage[warp];
    const int ns = n_raw < STAGE_CAP ? n_raw : STAGE_CAP;
    // ordered emission: rank by query offset, copies of a match dropped
    Match *dst = w.match_slots + slot_base(w, read);
    const int nu = emit_ranked(sm.stage[warp], ns, dst, w.cap, lane);
    const int n = nu + (n_raw - ns);                                    // distinct matches (+ what did not fit the stage)
    int n_out = n;
    if (sp.mum && n <= w.cap && n <= STAGE_CAP && nsurv == 0) {        // (with parked candidates k_mam_verify runs the sweep)
      // -mum: the sweep needs the matches in emission (query) order -> copy the ordered slots back into the
      // stage, let lane 0 run the by_ref sort + cleanMUMcand sweep, survivors go to the slots in by_ref order
      __syncwarp();
      for (int e = lane; e < nu; e += 32) sm.stage[warp][e] = dst[e];
      __syncwarp();
      if (lane == 0) {
        uint16_t ord[STAGE_CAP];
        n_out = mum_clean(sm.stage[warp], nu, ord, dst);
      }
      n_out = __shfl_sync(0xffffffffu, n_out, 0);
    }
    if (lane == 0) {
      w.match_cnt[read] = (uint32_t)n_out;
      if (w.surv) w.surv_cnt[read] = (uint8_t)nsurv;
      // more staged entries than the stage holds (copies of a saturated repeat family included): the dropped ones may
      // have been distinct matches, so this is always an overflow, never a silent count
      if (n > w.cap || n_raw > STAGE_CAP) { atomicAdd(&w.flags[FLAG_OVERFLOW], 1u); atomicMax(&w.flags[FLAG_MAXCNT], (uint32_t)(n_raw > STAGE_CAP ? n_raw : n)); }
    }
    __syncwarp();
  }
}

// K1a': the seed stage of the split search, written for instruction count.  One warp per read, lanes = anchors, no
// task lists and no per-byte staging:
//   (0) the read's aligned words are fetched once (coalesced); each lane lower-cases its word through the per-CTA table,
//       writes it to the lower-cased HBM copy k_mam_verify works on, and reduces it to ONE byte of 2-bit codes; an
//       8-lane OR builds the non-acgt bit mask.  The 2-bit stream lives in shared memory (38 bytes for a 150 bp read).
//   (1) lane a takes anchor x = a*s: one unaligned 64-bit window of the 2-bit stream holds, in order, the 4 bases
//       before the k-mer, the k-mer and the 4 bases after it -> k-mer code, both halves of the read's ext code;
//   (2) seed lookup [S[x], S[x+1]);
//   (3) the bucket entries' 2-byte ext codes are tested where they lie (first four entries of every lane at once, the
//       rare longer buckets in a warp-uniform loop) and the survivors are parked for k_mam_verify by ballot.
// Reads that need the exact machinery -- a non-acgt byte that occurs in the text, a bucket over BIG_BUCKET, more than
// SURV_CAP survivors -- are flagged in w.slow and redone by k_mam_search, which skips everything else.
constexpr int CODE_PAD = 4;                                   // code bytes (16 bases) in front of the stream
constexpr int CODE_BYTES = (MAXQ_FAST + 8) / 4 + CODE_PAD + 12;
struct SeedSmem {
  uint16_t lut[256];                                           // as SearchSmem::lut
  uint8_t code[WARPS][(CODE_BYTES + 3) & ~3];                  // byte CODE_PAD + i: 2-bit codes of buffer bytes 4i..4i+3, first base in the top bits
  uint32_t inv[WARPS][MAXQ_FAST / 32 + 12];                    // non-acgt mask in buffer coordinates (bit j + mis for base j)
};
// 8 bytes of the code stream starting at byte `bi`, as one big-endian number (first base in the top bits)
__device__ __forceinline__ uint64_t code_window(const uint8_t *code, int bi) {
  const uint32_t *a = reinterpret_cast<const uint32_t *>(code + (bi & ~3));
  const uint32_t w0 = __byte_perm(a[0], 0, 0x0123), w1 = __byte_perm(a[1], 0, 0x0123), w2 = __byte_perm(a[2], 0, 0x0123);
  const unsigned sh = (unsigned)(bi & 3) * 8u;
  return ((uint64_t)__funnelshift_l(w1, w0, sh) << 32) | __funnelshift_l(w2, w1, sh);
}
#ifndef SMASH_SEED_MINBLK
#define SMASH_SEED_MINBLK 6
#endif
__global__ void __launch_bounds__(THREADS, SMASH_SEED_MINBLK)
k_mam_seed(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  __shared__ __align__(16) SeedSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const uint8_t c = query_char((uint8_t)threadIdx.x, sp.nucleotides_only);
    const bool not_acgt = base_code(c) > 3;
    sm.lut[threadIdx.x] = (uint16_t)(c | (not_acgt ? 0x100 : 0) | (not_acgt && in_alpha(ix, c) ? 0x200 : 0));
  }
  __syncthreads();
  const int L = (int)sp.L, s = sp.s, k = sp.k;
  uint8_t *code = sm.code[warp];
  uint32_t *inv = sm.inv[warp];
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t read = (uint64_t)blockIdx.x * WARPS + warp; read < b.n_reads; read += warps_total) {
    const int64_t so = b.seq_off[read];
    const int q = (int)(b.seq_off[read + 1] - so);
    if (q > MAXQ_FAST) {                           // k_mam_search_long takes these (the host learns the length from the flag)
      if (lane == 0) { atomicMax(&w.flags[FLAG_LONGQ], (uint32_t)q); if (q > w.long_q) w.match_cnt[read] = 0; w.surv_cnt[read] = 0; w.slow[read] = 0; }
      continue;
    }
    // (0) words in, lower-cased copy out, 2-bit stream + non-acgt mask into shared memory
    const int mis = (int)(so & 3);
    const uint32_t *gw = reinterpret_cast<const uint32_t *>(b.seq + (so - mis));
    uint8_t *lc = w.lc + so + 32 * read + 16;
    uint32_t *lw = reinterpret_cast<uint32_t *>(lc - mis);
    const int nwords = (mis + q + 3) >> 2;
    unsigned oddbits = 0;
    int w0 = 0;
    for (; w0 < nwords; w0 += 32) {
      const int i = w0 + lane;
      unsigned nib = 0;
      if (i < nwords) {
        const uint32_t o = lut4(sm.lut, __ldg(gw + i), 4 * i - mis, q, nib, oddbits);
        lw[i] = o;
        code[CODE_PAD + i] = (uint8_t)code4(o);
        if (4 * i - mis < 0 || 4 * i - mis + 3 >= q) {        // edge words: bytes outside the read can start no k-mer
          for (int t = 0; t < 4; ++t) { const int j = 4 * i - mis + t; if (j < 0 || j >= q) nib |= 1u << t; }
        }
      }
      unsigned m = nib << (4 * (lane & 7));
      m |= __shfl_xor_sync(0xffffffffu, m, 1); m |= __shfl_xor_sync(0xffffffffu, m, 2); m |= __shfl_xor_sync(0xffffffffu, m, 4);
      if ((lane & 7) == 0) inv[i >> 3] = m;
    }
    if (lane < 2) inv[(w0 >> 3) + lane] = 0xffffffffu;        // spare words (kmer_invalid reads word + 1): nothing starts there
    if (lane < CODE_PAD) code[lane] = 0;
    if (lane < 12) code[CODE_PAD + nwords + lane] = 0;
    // pads of the lower-cased copy (core.cuh: 0xFE before, 0xFF after the read)
    if (lane < P_FRONT && lane - P_FRONT < -mis) lc[lane - P_FRONT] = 0xFE;
    { const int j = 4 * nwords - mis + lane; if (j < q + P_BACK) lc[j] = 0xFF; }
    __syncwarp();
    const bool odd = __any_sync(0xffffffffu, oddbits != 0);
    int nsurv = 0;
    bool slow = odd;
    if (q >= L && !odd) {
      const int n_anchor = (q - L + s - 1) / s + 1;          // anchors x = a*s cover starts 0..q-L
      for (int a0 = 0; a0 < n_anchor && !slow; a0 += 32) {
        const int a = a0 + lane, x = a * s;
        uint64_t lo = 0; int cnt = 0; uint32_t rext = 0, lvr = 0;
        const uint32_t *entries = ix.ext; int jb = 0;          // the bucket's ext codes (inside the seed line, or the flat table)
        if (a < n_anchor && !kmer_invalid(inv, x + mis, k)) {
          // (1) two windows of the 2-bit stream: bases x-8 .. x+k (8 before the k-mer + the k-mer), bases x+k .. x+k+6
          const int xb = x + mis + 4 * CODE_PAD - 8;           // stream position of base x-8
          const uint64_t win = code_window(code, xb >> 2) << (2 * (xb & 3));
          const uint32_t lcode = (uint32_t)(win >> 48);        // base x-8 in the top bits
          const uint64_t kc = (win >> (48 - 2 * k)) & ((1ull << (2 * k)) - 1ull);
          const int xr = xb + 8 + k;
          const uint32_t rcode = (uint32_t)((code_window(code, xr >> 2) << (2 * (xr & 3))) >> 52);   // 6 bases after the k-mer
          rext = rcode | (lcode << 12);
          // how many read bases immediately left of x are acgt (0..8; bases before the read's start are not)
          {
            const int p = x + mis;                             // mask bit of base x (buffer coordinates)
            uint32_t m8;
            if (p >= 8) { const int q0 = p - 8, wd = q0 >> 5, sh = q0 & 31; uint32_t m = inv[wd] >> sh; if (sh > 24) m |= inv[wd + 1] << (32 - sh); m8 = m & 0xffu; }
            else m8 = ((inv[0] << (8 - p)) | ((1u << (8 - p)) - 1u)) & 0xffu;
            lvr = m8 ? (uint32_t)__clz((int)(m8 << 24)) : 8u;  // bit 7 = base x-1
          }
          // (2) seed bucket (a sorted superset of the k-mer's suffix-array interval)
          const int shk = 2 * (ix.seed_k - k);
          if (ix.seed_blocked && shk == 0) {
            // one 128-byte line: block header (rank of the block's first suffix + 16 bucket sizes) and, behind it, the
            // ext codes of the block's first SEED_INLINE suffixes.  32-bit arithmetic throughout: the sizes are four
            // words of four bytes, a byte sum is one __vsadu4.
            const uint32_t *blk = reinterpret_cast<const uint32_t *>(ix.seed) + ((kc >> 4) << 5);
            const uint4 h = __ldg(reinterpret_cast<const uint4 *>(blk));
            const uint2 h2 = __ldg(reinterpret_cast<const uint2 *>(blk + 4));
            if (h.y >> 31) cnt = BIG_BUCKET + 1;               // a bucket of >= 255 suffixes in the block: exact path
            else {
              const unsigned bi = (unsigned)kc & 15u, wsel = bi >> 2, bsh = (bi & 3u) * 8u;
              const uint32_t below = (1u << bsh) - 1u;          // bytes of the selected word in front of this bucket
              const uint32_t cw = wsel == 0 ? h.z : wsel == 1 ? h.w : wsel == 2 ? h2.x : h2.y;
              jb = (int)(__vsadu4(cw & below, 0u) + (wsel > 0 ? __vsadu4(h.z, 0u) : 0u) + (wsel > 1 ? __vsadu4(h.w, 0u) : 0u) +
                         (wsel > 2 ? __vsadu4(h2.x, 0u) : 0u));
              const int c = (int)((cw >> bsh) & 0xffu);
              lo = ((uint64_t)h.x | ((uint64_t)(h.y & 0xffu) << 32)) + (uint64_t)jb;
              cnt = c < BIG_BUCKET + 1 ? c : BIG_BUCKET + 1;
              // the whole bucket inside the line (the usual case): its codes are read from there, else from the flat table
              entries = jb + c <= SEED_INLINE ? blk + 6 + jb : ix.ext + lo;
            }
          } else {
            lo = seed_at(ix, kc << shk);
            const uint64_t hi = seed_at(ix, (kc + 1) << shk);
            cnt = (int)(hi - lo < (uint64_t)(BIG_BUCKET + 1) ? hi - lo : (uint64_t)(BIG_BUCKET + 1));
            entries = ix.ext + lo;
          }
        }
        if (__any_sync(0xffffffffu, cnt > BIG_BUCKET)) { slow = true; break; }
        // (3) ext filter on the bucket entries: reach test + EXACT left extension (ownership); survivors parked in
        // (entry, lane) order with their left extension
        for (int j0 = 0; !slow && __any_sync(0xffffffffu, j0 < cnt); j0 += 4) {
          uint32_t e[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) e[t] = j0 + t < cnt ? __ldg(entries + j0 + t) : 0u;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            bool pass = false; uint32_t left = 0;
            if (j0 + t < cnt) {
              const uint32_t xo = e[t] ^ rext;
              const uint32_t r6 = ext_right_pairs(xo & 0xfffu), lm = ext_left_pairs((xo >> 12) & 0xffffu);
              const uint32_t lvt = e[t] >> 28, v = lvt < lvr ? lvt : lvr;
              left = lm < v ? lm : v;                          // exact when < 8: a mismatch, or a byte nothing here can match
              pass = left == 8 ? s > 8 : ((int)left < s && (r6 == 6 || (uint32_t)k + left + r6 >= sp.L));
            }
            const unsigned mask = __ballot_sync(0xffffffffu, pass);
            if (!mask) continue;
            const int np = __popc(mask);
            if (nsurv + np > SURV_CAP) { slow = true; break; }
            if (pass) w.surv[read * SURV_CAP + nsurv + __popc(mask & ((1u << lane) - 1u))] = ((uint64_t)x << 48) | ((uint64_t)left << 40) | (lo + (uint64_t)(j0 + t));
            nsurv += np;
          }
        }
      }
    }
    if (lane == 0) {
      w.slow[read] = slow ? 1 : 0;
      w.surv_cnt[read] = (uint8_t)(slow ? 0 : nsurv);
      w.match_cnt[read] = 0;
    }
    __syncwarp();
  }
}

// K1b: verification of the parked candidates.  k_mam_seed parks only candidates that OWN their diagonal, with the
// left extension it read off the ext code, so a read is left with a handful of them: 8 lanes per read, four reads per
// warp.  Lane = candidate: SA entry, the right extension along the diagonal (8-byte compares of the text against the
// lower-cased read copy in HBM), one byte of U.  The matches join the ones k_mam_search found on its exact paths, get
// rank-sorted by query offset (copies dropped, see emit_ranked) and are written in order.
constexpr int VG = 4;                                          // reads per warp
struct VerifySmem { Match stage[WARPS][VG][STAGE_CAP]; };
// emit_ranked for an 8-lane group (sl = lane & 7, the group's lanes are bits gshift..gshift+7 of a ballot); every lane of
// the warp must call it (groups without work pass ns = 0)
__device__ __forceinline__ int emit_ranked_group(const Match *stage, int ns, Match *dst, int cap, int sl, int gshift) {
  static_assert(STAGE_CAP <= 64, "one 64-bit mask covers the stage");
  uint64_t first = 0;                                          // bit e: entry e is the first with its query offset
  for (int j = 0; __any_sync(0xffffffffu, 8 * j < ns); ++j) {
    const int e = 8 * j + sl;
    bool fst = false;
    if (e < ns) {
      fst = true;
      const uint32_t qp = stage[e].qpos;
      for (int f = 0; f < e; ++f) if (stage[f].qpos == qp) { fst = false; break; }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, fst);
    first |= (uint64_t)((bal >> gshift) & 0xffu) << (8 * j);
  }
  for (int e = sl; e < ns; e += 8) {
    if (!((first >> e) & 1ull)) continue;
    const Match me = stage[e];
    int rank = 0;
    for (int f = 0; f < ns; ++f) rank += (((first >> f) & 1ull) && stage[f].qpos < me.qpos) ? 1 : 0;
    if (rank < cap) dst[rank] = me;
  }
  return __popcll(first);
}
#ifndef SMASH_VERIFY_MINBLK
#define SMASH_VERIFY_MINBLK 6
#endif
__global__ void __launch_bounds__(THREADS, SMASH_VERIFY_MINBLK)
k_mam_verify(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  __shared__ __align__(16) VerifySmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 3, sl = lane & 7, gshift = 8 * g;
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  const uint8_t *T = ix.text;
  Match *stage = sm.stage[warp][g];
  for (uint64_t r0 = ((uint64_t)blockIdx.x * WARPS + warp) * VG; r0 < b.n_reads; r0 += warps_total * VG) {
    const uint64_t read = r0 + (uint64_t)g;
    const int nsv = read < b.n_reads ? (int)w.surv_cnt[read] : 0;        // 0: k_mam_search finished this read (or none)
    if (!__any_sync(0xffffffffu, nsv > 0)) continue;
    int64_t so = 0; int q = 0; const uint8_t *P = nullptr; Match *dst = nullptr;
    int n_e_true = 0, n_e = 0;
    if (nsv) {
      so = b.seq_off[read];
      q = (int)(b.seq_off[read + 1] - so);
      P = w.lc + so + 32 * read + 16;
      dst = w.match_slots + slot_base(w, read);
      n_e_true = (int)w.match_cnt[read];                     // found by k_mam_search's exact paths (already in the slots)
      n_e = n_e_true < w.cap ? n_e_true : w.cap;
      if (n_e > STAGE_CAP) n_e = STAGE_CAP;
      for (int e = sl; e < n_e; e += 8) stage[e] = dst[e];
    }
    int n_new = 0;
    for (int base = 0; __any_sync(0xffffffffu, base < nsv); base += 8) {
      const int idx = base + sl;
      Match m; int r = 0;
      if (idx < nsv) {
        const uint64_t e = w.surv[read * SURV_CAP + idx];
        const int x = (int)(e >> 48);
        const uint64_t c = sa_at(ix, e & 0xffffffffffull);
        int left = (int)((e >> 40) & 0xffu);                 // from the ext code: exact below 8
        if (left >= 8) left = match_left(T, (int64_t)c, P, x, sp.s);
        if (left < sp.s) {                                   // this anchor owns the diagonal (candidate_check, core.cuh)
          const int right = match_right(T, (int64_t)c, P, x, q - x);
          const uint32_t len = (uint32_t)(left + right);
          if (right >= sp.k && len >= sp.L && len >= 2) {
            const uint64_t ref = c - (uint64_t)left;
            const uint8_t u = ix.uniq[ref];
            if (u == 255 && len >= 255) r = exact_start<false>(ix, P, q, x - left, sp.L, &m) ? 1 : 0;
            else if (len >= u) { m.ref = ref; m.qpos = (uint32_t)(x - left); m.len = len; r = 1; }
          }
        }
      }
      const unsigned pass = (__ballot_sync(0xffffffffu, r > 0) >> gshift) & 0xffu;
      if (r > 0) { const int pos = n_e + n_new + __popc(pass & ((1u << sl) - 1u)); if (pos < STAGE_CAP) stage[pos] = m; }
      n_new += __popc(pass);
    }
    __syncwarp();
    const int n_raw = n_e_true + n_new;
    const int ns = n_e + n_new < STAGE_CAP ? n_e + n_new : STAGE_CAP;
    const int nu = emit_ranked_group(stage, ns, dst, w.cap, sl, gshift);
    const int n = nu + (n_raw - ns);                                    // distinct matches (+ what did not fit the stage)
    int n_out = n;
    if (sp.mum) {                                                       // -mum: cleanMUMcand sweep on the ordered list
      __syncwarp();
      const bool sweep = nsv && n <= w.cap && n <= STAGE_CAP;
      if (sweep) for (int e = sl; e < nu; e += 8) stage[e] = dst[e];
      __syncwarp();
      if (sweep && sl == 0) { uint16_t ord[STAGE_CAP]; n_out = mum_clean(stage, nu, ord, dst); }
    }
    if (nsv && sl == 0) {
      w.match_cnt[read] = (uint32_t)n_out;
      // more staged entries than the stage holds (copies of a saturated repeat family included): the dropped ones may
      // have been distinct matches, so this is always an overflow, never a silent count
      if (n > w.cap || n_raw > STAGE_CAP) { atomicAdd(&w.flags[FLAG_OVERFLOW], 1u); atomicMax(&w.flags[FLAG_MAXCNT], (uint32_t)(n_raw > STAGE_CAP ? n_raw : n)); }
    }
    __syncwarp();
  }
}

// Reads longer than the shared-memory staging buffer: staged (lower-cased, padded) in this warp's HBM
// scratch and searched with the exact per-start path.  Launched only when the batch has such reads.
struct LongSmem { Match stage[WARPS][STAGE_CAP]; int nstage[WARPS]; };
__global__ void __launch_bounds__(THREADS)
k_mam_search_long(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  __shared__ __align__(16) LongSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t read = (uint64_t)blockIdx.x * WARPS + warp; read < b.n_reads; read += warps_total) {
    const int64_t so = b.seq_off[read];
    const int q = (int)(b.seq_off[read + 1] - so);
    if (q <= MAXQ_FAST) continue;
    if (lane == 0) sm.nstage[warp] = 0;
    __syncwarp();
    if (!w.long_scratch || q > w.long_q) {
      if (lane == 0) { atomicAdd(&w.flags[FLAG_LONGREAD], 1u); w.match_cnt[read] = 0; }
      continue;
    }
    uint8_t *gbuf = w.long_scratch + ((uint64_t)blockIdx.x * WARPS + warp) * (uint64_t)(w.long_q + P_FRONT + P_BACK + 8);
    stage_read(ix, b.seq + so, q, sp.nucleotides_only, gbuf, lane);
    __threadfence_block();
    const uint8_t *P = gbuf + P_FRONT;
    for (int p = lane; p + (int)sp.L <= q; p += 32) {
      Match m;
      if (exact_start(ix, P, q, p, sp.L, &m)) {
        const int slot = atomicAdd(&sm.nstage[warp], 1);
        if (slot < STAGE_CAP) sm.stage[warp][slot] = m;
      }
    }
    __syncwarp();
    const int n_raw = sm.nstage[warp];
    const int ns = n_raw < STAGE_CAP ? n_raw : STAGE_CAP;
    Match *dst = w.match_slots + slot_base(w, read);
    const int nu = emit_ranked(sm.stage[warp], ns, dst, w.cap, lane);
    const int n = nu + (n_raw - ns);                                    // distinct matches (+ what did not fit the stage)
    int n_out = n;
    if (sp.mum && n <= w.cap && n <= STAGE_CAP) {
      __syncwarp();
      for (int e = lane; e < nu; e += 32) sm.stage[warp][e] = dst[e];
      __syncwarp();
      if (lane == 0) { uint16_t ord[STAGE_CAP]; n_out = mum_clean(sm.stage[warp], nu, ord, dst); }
      n_out = __shfl_sync(0xffffffffu, n_out, 0);
    }
    if (lane == 0) {
      w.match_cnt[read] = (uint32_t)n_out;
      // more staged entries than the stage holds (copies of a saturated repeat family included): the dropped ones may
      // have been distinct matches, so this is always an overflow, never a silent count
      if (n > w.cap || n_raw > STAGE_CAP) { atomicAdd(&w.flags[FLAG_OVERFLOW], 1u); atomicMax(&w.flags[FLAG_MAXCNT], (uint32_t)(n_raw > STAGE_CAP ? n_raw : n)); }
    }
    __syncwarp();
  }
}

// K1c: any number of matches per read.  The anchor kernels stage at most STAGE_CAP matches of a read in shared memory; a
// range holding a read with more (the reference has no such limit: a few-thousand-base read made of short unique pieces)
// is redone here with the exact per-start path and CSR slots, as MEM mode does: a COUNT pass, the slot offsets, a WRITE
// pass.  One warp per read, lanes = query starts p (32 at a time, in increasing order), so a ballot prefix writes the
// matches in query order without any staging.  -mum: the raw matches go to the read's scratch slots and lane 0 runs the
// cleanMUMcand sweep from there into the match slots.
template <bool WRITE>
__global__ void __launch_bounds__(THREADS)
k_mam_exact(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp, uint32_t *__restrict__ cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  uint8_t *gbuf = w.long_scratch + ((uint64_t)blockIdx.x * WARPS + warp) * (uint64_t)(w.long_q + P_FRONT + P_BACK + 8);
  for (uint64_t read = (uint64_t)blockIdx.x * WARPS + warp; read < b.n_reads; read += warps_total) {
    const int64_t so = b.seq_off[read];
    const int q = (int)(b.seq_off[read + 1] - so);
    if (q > w.long_q) {                                        // cannot happen: the host sized the scratch for the longest read
      if (lane == 0) { atomicAdd(&w.flags[FLAG_LONGREAD], 1u); if (WRITE) w.match_cnt[read] = 0; else cnt[read] = 0; }
      continue;
    }
    __syncwarp();
    stage_read(ix, b.seq + so, q, sp.nucleotides_only, gbuf, lane);
    __threadfence_block();
    const uint8_t *P = gbuf + P_FRONT;
    Match *dst = nullptr;
    if (WRITE) dst = sp.mum ? reinterpret_cast<Match *>(w.aln_scratch + slot_base(w, read)) : w.match_slots + slot_base(w, read);
    int n = 0;
    for (int p0 = 0; p0 + (int)sp.L <= q; p0 += 32) {
      const int p = p0 + lane;
      Match m;
      const bool ok = p + (int)sp.L <= q && exact_start(ix, P, q, p, sp.L, &m);
      const unsigned mask = __ballot_sync(0xffffffffu, ok);
      if (WRITE && ok) dst[n + __popc(mask & ((1u << lane) - 1u))] = m;
      n += __popc(mask);
    }
    if (WRITE) {
      __syncwarp();
      if (lane == 0) {
        if (sp.mum) n = mum_clean(dst, n, w.ord_scratch + slot_base(w, read), w.match_slots + slot_base(w, read));
        w.match_cnt[read] = (uint32_t)n;
      }
    } else if (lane == 0) {
      cnt[read] = (uint32_t)n;
      if (n > 65535) atomicAdd(&w.flags[FLAG_OVERFLOW], 1u);   // 16-bit per-read record fields
    }
  }
}
int launch_mam_exact(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, uint32_t *cnt, cudaStream_t st) {
  if (!b.n_reads) return 0;
  if (cnt) k_mam_exact<false><<<grid_for_warps(b.n_reads, 8), THREADS, 0, st>>>(ix, b, w, p, cnt);
  else k_mam_exact<true><<<grid_for_warps(b.n_reads, 8), THREADS, 0, st>>>(ix, b, w, p, nullptr);
  return 1;
}

int launch_mam_search(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st) {
  if (!b.n_reads) return 0;
  if (w.slow) {                                    // split search with the lean seed stage; k_mam_search redoes the flagged reads
    k_mam_seed<<<grid_for_warps(b.n_reads, SMASH_SEED_MINBLK), THREADS, 0, st>>>(ix, b, w, p);
    k_mam_search<<<grid_for_warps(b.n_reads, 8), THREADS, 0, st>>>(ix, b, w, p);
    return 2;
  }
  k_mam_search<<<grid_for_warps(b.n_reads, 8), THREADS, 0, st>>>(ix, b, w, p);
  return 1;
}
// after launch_mam_search: extension of the parked candidates (split search) and the long-read kernel
int launch_mam_verify(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st) {
  if (!b.n_reads) return 0;
  int nl = 0;
  if (w.surv) { k_mam_verify<<<grid_for_warps((b.n_reads + VG - 1) / VG, SMASH_VERIFY_MINBLK), THREADS, 0, st>>>(ix, b, w, p); ++nl; }
  if (w.long_q > MAXQ_FAST) { k_mam_search_long<<<grid_for_warps(b.n_reads, 8), THREADS, 0, st>>>(ix, b, w, p); ++nl; }
  return nl;
}

// ------------------------------------------------------------------ K3: records
//
// k_rec_build   one THREAD per read: resolve / erase / sort / merge / HI order (query.cpp:68-97,
//               231-320) with the per-read scratch in thread-local arrays -- 32 reads advance per
//               warp instead of one lane working while 31 wait.  (Reads with more than LOCAL_CAP
//               matches only exist after a slot-capacity rerun; they take k_rec_build_serial.)
// scan          nrec -> rec_base: every record gets a flat index
// k_rec_xe      one warp per read, 8 lanes per record: XE (query.cpp:270-274) straight from the
//               read and text words in HBM, L/R mappability of every '=' block, flat record -> read map
constexpr int LOCAL_CAP = 24;

__global__ void __launch_bounds__(128)
k_rec_build(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  for (uint64_t read = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; read < b.n_reads; read += (uint64_t)gridDim.x * blockDim.x) {
    const int q = (int)(b.seq_off[read + 1] - b.seq_off[read]);
    int n_in = (int)w.match_cnt[read];
    if (n_in > slot_cap(w, read)) n_in = slot_cap(w, read);
    if (n_in > LOCAL_CAP) n_in = LOCAL_CAP;
    Aln aln[LOCAL_CAP]; uint16_t ord[LOCAL_CAP];
    ReadSum sum;
    const int n_rec = build_records(ix, w.match_slots + slot_base(w, read), n_in, q, sp.nomap, aln, ord,
                                    w.item_slots + slot_base(w, read), w.rec_slots + slot_base(w, read), &sum);
    w.sums[read] = sum;
    w.nrec[read] = (uint32_t)n_rec;
  }
}

struct RecSmem {
  Aln aln[WARPS][SCR_CAP];
  uint16_t ord[WARPS][SCR_CAP];
};
__global__ void __launch_bounds__(THREADS)
k_rec_build_serial(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  __shared__ __align__(16) RecSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t read = (uint64_t)blockIdx.x * WARPS + warp; read < b.n_reads; read += warps_total) {
    const int q = (int)(b.seq_off[read + 1] - b.seq_off[read]);
    int n_in = (int)w.match_cnt[read];
    if (n_in > slot_cap(w, read)) n_in = slot_cap(w, read);
    if (n_in > SCR_CAP) n_in = SCR_CAP;
    if (lane == 0) {
      const int n_rec = build_records(ix, w.match_slots + slot_base(w, read), n_in, q, sp.nomap, sm.aln[warp], sm.ord[warp],
                                      w.item_slots + slot_base(w, read), w.rec_slots + slot_base(w, read), &w.sums[read]);
      w.nrec[read] = (uint32_t)n_rec;
    }
    __syncwarp();
  }
}

// MEM mode: any number of matches per read, scratch in HBM next to the read's slots (lane 0 works;
// -maxmatch is the secondary mode and its per-read record logic is inherently serial).
__global__ void __launch_bounds__(THREADS)
k_rec_build_big(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t read = (uint64_t)blockIdx.x * WARPS + warp; read < b.n_reads; read += warps_total) {
    if (lane == 0) {
      const int q = (int)(b.seq_off[read + 1] - b.seq_off[read]);
      const uint64_t base = slot_base(w, read);
      int n_in = (int)w.match_cnt[read];
      if (n_in > slot_cap(w, read) - 1) n_in = slot_cap(w, read) - 1;
      const int n_rec = build_records(ix, w.match_slots + base, n_in, q, sp.nomap, w.aln_scratch + base, w.ord_scratch + base,
                                      w.item_slots + base, w.rec_slots + base, &w.sums[read]);
      w.nrec[read] = (uint32_t)n_rec;
    }
  }
}

// 8 bytes of the read at offset j (unaligned) with NewQuery::extend's tolower applied (A-Z only)
__device__ __forceinline__ uint64_t read8_lower(const uint8_t *__restrict__ seq, int j) {
  const uint8_t *p = seq + j;
  const uint64_t *a = reinterpret_cast<const uint64_t *>(p - ((uintptr_t)p & 7));
  const unsigned sh = (unsigned)((uintptr_t)p & 7) * 8u;
  uint64_t x = __ldg(a);
  if (sh) x = (x >> sh) | (__ldg(a + 1) << (64u - sh));
  const uint64_t h = x & 0x7f7f7f7f7f7f7f7fULL;
  const uint64_t ge_A = h + 0x3f3f3f3f3f3f3f3fULL, gt_Z = h + 0x2525252525252525ULL;
  const uint64_t upper = ge_A & ~gt_Z & ~x & 0x8080808080808080ULL;
  return x | (upper >> 2);
}
__device__ __forceinline__ int xe_word_g(const DevIndex &ix, const uint8_t *__restrict__ seq, int q, int64_t rcpos, int j0, int nuc) {
  const int64_t rp = rcpos + j0;
  if (!nuc && rp >= 0 && rp + 8 <= (int64_t)ix.N && j0 + 8 <= q) {
    const uint64_t d = text8(ix.text, rp) ^ read8_lower(seq, j0);
    uint64_t t = (d & 0x7f7f7f7f7f7f7f7fULL) + 0x7f7f7f7f7f7f7f7fULL;
    t = ~(t | d | 0x7f7f7f7f7f7f7f7fULL);
    return __popcll(t);
  }
  int cnt = 0;
  for (int j = j0; j < j0 + 8 && j < q; ++j) {
    const int64_t p = rcpos + j;
    if (p >= 0 && p < (int64_t)ix.N && ix.text[p] == query_char(seq[j], nuc)) ++cnt;
  }
  return cnt;
}

// 8 bytes of a lower-cased staged read (w.lc: 16 pad bytes either side) at offset j, unaligned
__device__ __forceinline__ uint64_t lc8(const uint8_t *__restrict__ P, int j) {
  const uint8_t *p = P + j;
  const uint64_t *a = reinterpret_cast<const uint64_t *>(p - ((uintptr_t)p & 7));
  const unsigned sh = (unsigned)((uintptr_t)p & 7) * 8u;
  const uint64_t lo = a[0];
  if (sh == 0) return lo;
  return (lo >> sh) | (a[1] << (64u - sh));
}
__device__ __forceinline__ int zero_bytes(uint64_t d) {      // number of zero bytes of d
  uint64_t t = (d & 0x7f7f7f7f7f7f7f7fULL) + 0x7f7f7f7f7f7f7f7fULL;
  t = ~(t | d | 0x7f7f7f7f7f7f7f7fULL);
  return __popcll(t);
}

// K3b, flat: 8 lanes per RECORD, four records per warp pass, over the batch's flat record list (rec_read was filled by
// the rec_base scan).  XE (query.cpp:270-274) = matching characters of the WHOLE read along the record's diagonal:
// lane t compares the 8-byte words t, t+8, t+16, .. of the lower-cased read copy the search stage left in HBM (its
// 0xFE/0xFF pads and the zero pads of the text never match anything, so neither end needs a branch) against the text.
// Then lanes = the record's '=' blocks: L/R of each from map.bin (map_lr), kept in the Item for the size / emit / tail
// kernels, and the tagger's range check.
#ifndef SMASH_XE_MINBLK
#define SMASH_XE_MINBLK 6
#endif
__global__ void __launch_bounds__(THREADS, SMASH_XE_MINBLK)
k_rec_xe(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 3, sl = lane & 7;
  const unsigned gmask = 0xffu << (8 * sub);
  const uint64_t n_records = w.rec_base[b.n_reads];
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t f0 = ((uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * 4; f0 < n_records; f0 += warps_total * 4) {
    const uint64_t f = f0 + (uint64_t)sub;
    const bool have = f < n_records;
    uint64_t read = 0; int r = 0; bool mapped = false;
    if (have) { read = w.rec_read[f]; r = (int)(f - w.rec_base[read]); mapped = !w.sums[read].unmapped; }
    int cnt = 0;
    Rec *rec = nullptr; Item *items = nullptr;
    int64_t pos = 0; uint32_t si = 0; int item_begin = 0, item_cnt = 0;
    if (mapped) {
      const uint64_t sb = slot_base(w, read);
      rec = w.rec_slots + sb + r; items = w.item_slots + sb;
      const int64_t rcpos = rec->rcpos;
      pos = rec->pos; si = rec->si; item_begin = rec->item_begin; item_cnt = rec->item_cnt;
      const int64_t so = b.seq_off[read];
      const int q = (int)(b.seq_off[read + 1] - so);
      if (w.lc && q <= MAXQ_FAST) {
        const uint8_t *P = w.lc + so + 32 * read + 16;
        for (int j0 = sl * 8; j0 < q; j0 += 64) {
          const int64_t rp = rcpos + j0;
          if (rp >= -(int64_t)(TEXT_PAD - 8) && rp <= (int64_t)ix.N + (TEXT_PAD - 16)) cnt += zero_bytes(text8(ix.text, rp) ^ lc8(P, j0));
        }
      } else {
        const uint8_t *seq = b.seq + so;
        for (int j0 = sl * 8; j0 < q; j0 += 64) cnt += xe_word_g(ix, seq, q, rcpos, j0, sp.nucleotides_only);
      }
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    if (mapped && sl == 0) rec->xe = (uint16_t)cnt;
    if (ix.mapbody) {                                        // L/R of every '=' block + the tagger's range check
      bool bad = false;
      if (mapped) {
        for (int u = sl; u < item_cnt; u += 8) {
          Item it = items[item_begin + u];
          int L, R;
          bad |= !map_lr(ix, si >> 1, pos, it.prefix, it.len, &L, &R);
          it.L = (uint8_t)L; it.R = (uint8_t)R;
          items[item_begin + u] = it;
          if (u == 0) { rec->L0 = (uint8_t)L; rec->R0 = (uint8_t)R; }
        }
      }
      const unsigned bm = __ballot_sync(0xffffffffu, bad) & gmask;
      if (bm && sl == 0) {
        // mappability_tag.cpp:107-113 throws unless the chromosome is _gl000*/chrM
        const char *nm = ix.descr + ix.descr_off[si];
        const int nl = ix.descr_off[si + 1] - ix.descr_off[si];
        bool small = false;
        for (int i = 0; i + 4 <= nl; ++i) if (nm[i] == 'c' && nm[i + 1] == 'h' && nm[i + 2] == 'r' && nm[i + 3] == 'M') small = true;
        for (int i = 0; i + 6 <= nl; ++i) if (nm[i] == '_' && nm[i + 1] == 'g' && nm[i + 2] == 'l' && nm[i + 3] == '0' && nm[i + 4] == '0' && nm[i + 5] == '0') small = true;
        if (!small) atomicAdd(&w.flags[FLAG_MAPERR], 1u);
      }
    }
  }
}

static int exclusive_scan_u32(const uint32_t *in, uint64_t n, uint64_t *blk, uint64_t *out, cudaStream_t st, uint32_t *fill = nullptr);
int launch_records(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st) {
  if (!b.n_reads) return 0;
  if (w.slot_off) {
    k_rec_build_big<<<grid_for_warps(b.n_reads, 8), THREADS, 0, st>>>(ix, b, w, p);
  } else if (w.cap <= LOCAL_CAP) {
    const uint64_t need = (b.n_reads + 127) / 128, cap = (uint64_t)sm_count() * SMASH_RECB_CTAS;
    k_rec_build<<<(unsigned)(need < cap ? need : cap), 128, 0, st>>>(ix, b, w, p);
  } else {
    k_rec_build_serial<<<grid_for_warps(b.n_reads, 6), THREADS, 0, st>>>(ix, b, w, p);
  }
  int n = 1 + exclusive_scan_u32(w.nrec, b.n_reads, w.blk_sums, w.rec_base, st, w.rec_read);    // + flat record -> read map
  k_rec_xe<<<grid_for_warps(b.n_reads, SMASH_XE_CTAS), THREADS, 0, st>>>(ix, b, w, p);
  return n + 1;
}

// ------------------------------------------------------------------ K4a: sizes + scan

// flag + mate view of one read (set_mate, query.cpp:421-434)
__device__ __forceinline__ void read_mate(const BatchDev &b, const WorkDev &w, uint64_t read,
                                          uint16_t *flag, MateView *mv) {
  const ReadSum me = w.sums[read];
  const uint16_t mine = (uint16_t)(b.read_flag[read] | (me.unmapped ? 4 : 0));
  const uint64_t other = read ^ 1ull;
  if (other < b.n_reads) {
    const ReadSum ot = w.sums[other];
    const uint16_t of = (uint16_t)(b.read_flag[other] | (ot.unmapped ? 4 : 0));
    mate_view(mine, me, of, &ot, (read & 1ull) == 0, mv, flag);
  } else {
    mate_view(mine, me, 0, nullptr, true, mv, flag);
  }
}

// One THREAD per record (flat index): exact byte length of its SAM line.
__global__ void __launch_bounds__(128)
k_sizes(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp) {
  const uint64_t n_records = w.rec_base[b.n_reads];
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_records; f += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t read = w.rec_read[f];
    const int r = (int)(f - w.rec_base[read]);
    const ReadSum me = w.sums[read];
    uint16_t flag; MateView mv;
    read_mate(b, w, read, &flag, &mv);
    const int q = (int)(b.seq_off[read + 1] - b.seq_off[read]);
    const int name_len = (int)(b.name_off[read + 1] - b.name_off[read]);
    const int opt_len = b.opt ? (int)(b.opt_off[read + 1] - b.opt_off[read]) : 0;
    const Item *items = w.item_slots + slot_base(w, read);
    Rec *recs = w.rec_slots + slot_base(w, read);
    CountSink cs;
    put_head(cs, ix, (const char *)nullptr, name_len, flag, me.unmapped, recs, r, me.n_rec, items, mv);
    const uint32_t seq_at = cs.n;                                  // offset of the SEQ column in the line
    put_tags(cs, ix, me.unmapped, recs, r, me.n_rec, items);
    const uint32_t before_lr = cs.n;
    if (sp.tag_mappability && !me.unmapped) put_lr_tags(cs, ix, recs[r], items);
    recs[r].seq_off = (uint16_t)seq_at; recs[r].lr_len = (uint16_t)(cs.n - before_lr);
    w.rec_bytes[f] = cs.n + 2u * (uint32_t)q + 1u /*tab between SEQ and QUAL*/ + (uint32_t)opt_len + 1u /*\n*/;
    if (w.cmp_bytes) w.cmp_bytes[f] = cs.n - (uint32_t)name_len + 1u;      // compact transport: head + tags + lr-tags + \n
  }
}

constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_TILE = SCAN_ITEMS * SCAN_BLOCK;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total) {
  __shared__ uint64_t wsum[SCAN_BLOCK / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
  for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t s = lane < SCAN_BLOCK / 32 ? wsum[lane] : 0, si = s;
    for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
    if (lane < SCAN_BLOCK / 32) wsum[lane] = si - s;
    if (lane == 31) *total = si;
  }
  __syncthreads();
  const uint64_t r = wsum[warp] + inc - v;
  __syncthreads();
  return r;
}

__global__ void k_scan_tiles(const uint32_t *__restrict__ in, uint64_t n, const uint64_t *__restrict__ n_dev, uint64_t *__restrict__ blk) {
  __shared__ uint64_t tot;
  if (n_dev) { n = *n_dev; if ((uint64_t)blockIdx.x * SCAN_TILE >= n) return; }
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint64_t s = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += in[base + i];
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) blk[blockIdx.x] = tot;
}
__global__ void k_scan_top(uint64_t *blk, uint64_t n_blk, const uint64_t *__restrict__ n_dev) {
  __shared__ uint64_t tot;
  __shared__ uint64_t carry;
  if (n_dev) n_blk = (*n_dev + SCAN_TILE - 1) / SCAN_TILE;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t b0 = 0; b0 < n_blk; b0 += SCAN_BLOCK) {
    const uint64_t i = b0 + threadIdx.x;
    const uint64_t v = i < n_blk ? blk[i] : 0;
    const uint64_t ex = block_exclusive_scan(v, &tot);
    if (i < n_blk) blk[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) blk[n_blk] = carry;
}
// fill != null: out[i] .. out[i] + in[i] of `fill` get the value i (the flat record -> read map of the rec_base scan)
__global__ void k_scan_apply(const uint32_t *__restrict__ in, uint64_t n, const uint64_t *__restrict__ n_dev,
                             const uint64_t *__restrict__ blk, uint64_t *__restrict__ out, uint64_t *__restrict__ total_out,
                             uint32_t *__restrict__ fill) {
  __shared__ uint64_t tot;
  uint64_t n_blk = gridDim.x;
  if (n_dev) {
    n = *n_dev; n_blk = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) { out[0] = 0; if (total_out) *total_out = 0; }
    if (blockIdx.x >= n_blk) return;
  }
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS]; uint64_t s = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
  uint64_t ex = block_exclusive_scan(s, &tot) + blk[blockIdx.x];
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) {
      out[base + i] = ex;
      if (fill) for (uint32_t j = 0; j < v[i]; ++j) fill[ex + j] = (uint32_t)(base + i);
    }
    ex += v[i];
  }
  if (blockIdx.x == n_blk - 1 && threadIdx.x == 0) { out[n] = blk[n_blk]; if (total_out) *total_out = blk[n_blk]; }
}
static int exclusive_scan_u32(const uint32_t *in, uint64_t n, uint64_t *blk, uint64_t *out, cudaStream_t st, uint32_t *fill) {
  const uint64_t n_blk = (n + SCAN_TILE - 1) / SCAN_TILE;
  k_scan_tiles<<<(unsigned)n_blk, SCAN_BLOCK, 0, st>>>(in, n, nullptr, blk);
  k_scan_top<<<1, SCAN_BLOCK, 0, st>>>(blk, n_blk, nullptr);
  k_scan_apply<<<(unsigned)n_blk, SCAN_BLOCK, 0, st>>>(in, n, nullptr, blk, out, nullptr, fill);
  return 3;
}
// same, but the element count lives in device memory (*n_dev <= n_bound): tiles past it exit at once
static int exclusive_scan_u32_devn(const uint32_t *in, uint64_t n_bound, const uint64_t *n_dev, uint64_t *blk, uint64_t *out,
                                   uint64_t *total_out, cudaStream_t st) {
  const uint64_t n_blk = (n_bound + SCAN_TILE - 1) / SCAN_TILE;
  k_scan_tiles<<<(unsigned)n_blk, SCAN_BLOCK, 0, st>>>(in, 0, n_dev, blk);
  k_scan_top<<<1, SCAN_BLOCK, 0, st>>>(blk, 0, n_dev);
  k_scan_apply<<<(unsigned)n_blk, SCAN_BLOCK, 0, st>>>(in, 0, n_dev, blk, out, total_out, nullptr);
  return 3;
}

// The three small results the host needs between the scan and the emit kernels (SAM bytes, record
// count, flag counters) go to MAPPED pinned host memory with plain stores: a copy-engine transfer
// would queue behind the other slot's in-flight SAM download and stall this slot for its whole length.
__global__ void k_publish(uint64_t *__restrict__ host_small, const uint64_t *__restrict__ sam_total,
                          const uint64_t *__restrict__ rec_total, const uint32_t *__restrict__ flags) {
  if (threadIdx.x == 0) { host_small[0] = sam_total[0]; host_small[8] = *rec_total; host_small[9] = sam_total[1]; }
  if (threadIdx.x < N_FLAGS) ((uint32_t *)(host_small + 1))[threadIdx.x] = flags[threadIdx.x];
  __threadfence_system();
}
int launch_publish(uint64_t *host_small, const uint64_t *sam_total, const uint64_t *rec_total, const uint32_t *flags, cudaStream_t st) {
  k_publish<<<1, 32, 0, st>>>(host_small, sam_total, rec_total, flags);
  return 1;
}

int launch_sizes_scan(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st) {
  if (!b.n_reads) return 0;
  k_sizes<<<sm_count() * SMASH_SIZES_CTAS, 128, 0, st>>>(ix, b, w, p);
  // rec_off = exclusive scan of rec_bytes over the (device-side) record count; total -> sam_total[0]
  int n = 1 + exclusive_scan_u32_devn(w.rec_bytes, w.slots_total, w.rec_base + b.n_reads, w.blk_sums2, w.rec_off,
                                      w.sam_total, st);
  if (w.cmp_bytes)                                             // compact transport: the same for the compact text
    n += exclusive_scan_u32_devn(w.cmp_bytes, w.slots_total, w.rec_base + b.n_reads, w.blk_sums2, w.cmp_off, w.sam_total + 1, st);
  return n;
}

// ------------------------------------------------------------------ K4b: emit
//
// Two kernels write every record straight into its final place (rec_off) in the SAM buffer:
//  k_emit_text  the VARIABLE text (columns 2-9, tags, L/R tags, newline): ONE THREAD PER RECORD formats
//               it through a WordSink -- characters are gathered in a register and leave as aligned
//               8-byte stores -- so 32 records are formatted at once, no shared memory is needed and
//               the kernel runs at full occupancy.
//  k_emit_copy  the BULK bytes (name, SEQ, QUAL, optional fields): one warp per READ loads them once
//               into shared memory and streams them into each of the read's records (reverse-
//               complemented for reverse-strand records) with coalesced byte stores.
#ifndef SMASH_TEXT_MINBLK
#define SMASH_TEXT_MINBLK 1
#endif
__global__ void __launch_bounds__(128, SMASH_TEXT_MINBLK)
k_emit_text(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp, uint64_t n_records) {
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_records; f += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t read = w.rec_read[f];
    const int hi = (int)(f - w.rec_base[read]);
    const ReadSum me = w.sums[read];
    uint16_t flag; MateView mv;
    read_mate(b, w, read, &flag, &mv);
    const Item *items = w.item_slots + slot_base(w, read);
    const Rec *recs = w.rec_slots + slot_base(w, read);
    const int name_len = (int)(b.name_off[read + 1] - b.name_off[read]);
    const int q = (int)(b.seq_off[read + 1] - b.seq_off[read]);
    const int opt_len = b.opt ? (int)(b.opt_off[read + 1] - b.opt_off[read]) : 0;
    char *out = w.sam + w.rec_off[f];
    WordSink hs(out + name_len);
    put_head(hs, ix, (const char *)nullptr, 0, flag, me.unmapped, recs, hi, me.n_rec, items, mv);
    hs.finish();
    WordSink ts(out + name_len + hs.n + 2 * q + 1);
    put_tags(ts, ix, me.unmapped, recs, hi, me.n_rec, items);
    if (opt_len) {                                           // the optional fields sit between the tags and the L/R tags
      ts.finish();
      ts = WordSink(out + name_len + hs.n + 2 * q + 1 + ts.n + opt_len);
    }
    if (sp.tag_mappability && !me.unmapped) put_lr_tags(ts, ix, recs[hi], items);
    ts.ch('\n');
    ts.finish();
  }
}

// Compact transport (compact.h): the same formatting, but head, tags and L/R tags of a record form ONE contiguous run
// in the range's compact text, followed by the newline; the bytes the caller already has (name, SEQ, QUAL, optional
// fields) are not written at all -- the host puts the line together (expand.cpp) -- so k_emit_copy is not launched.
__global__ void __launch_bounds__(128, SMASH_TEXT_MINBLK)
k_emit_compact(DevIndex ix, BatchDev b, WorkDev w, SearchParams sp, uint64_t n_records) {
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_records; f += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t read = w.rec_read[f];
    const int hi = (int)(f - w.rec_base[read]);
    const ReadSum me = w.sums[read];
    uint16_t flag; MateView mv;
    read_mate(b, w, read, &flag, &mv);
    const Item *items = w.item_slots + slot_base(w, read);
    const Rec *recs = w.rec_slots + slot_base(w, read);
    const uint64_t co = w.cmp_off[f];
    WordSink s(w.cmp + co);
    put_head(s, ix, (const char *)nullptr, 0, flag, me.unmapped, recs, hi, me.n_rec, items, mv);
    const uint32_t head_len = s.n;
    put_tags(s, ix, me.unmapped, recs, hi, me.n_rec, items);
    const uint32_t tags_len = s.n - head_len;
    if (sp.tag_mappability && !me.unmapped) put_lr_tags(s, ix, recs[hi], items);
    s.ch('\n');
    s.finish();
    const uint32_t lr_len = s.n - head_len - tags_len;
    const bool rc = recs[hi].rc && !me.unmapped;
    uint4 *m = reinterpret_cast<uint4 *>(w.cmeta + f);
    const uint64_t so = w.sam_base + w.rec_off[f];
    m[0] = make_uint4((uint32_t)so, (uint32_t)(so >> 32), (uint32_t)co, (uint32_t)read);
    m[1] = make_uint4(head_len, tags_len, lr_len | (rc ? 0x80000000u : 0u), 0u);
  }
}

// The bulk of a record is ONE contiguous run "SEQ \t QUAL" (2q+1 bytes).  The warp keeps that run (and
// its reverse-complemented twin) in shared memory and streams it into every record of the read as
// aligned 16-byte stores: lanes = 16-byte chunks of the destination, each assembled from five shared
// words by funnel shifts; the <16 B before the first and after the last aligned chunk go out as one
// predicated byte store (lanes 0-15 head, 16-31 tail).
#ifndef SMASH_COPY_MAXQ
#define SMASH_COPY_MAXQ MAXQ_FAST
#endif
constexpr int COPY_MAXQ = SMASH_COPY_MAXQ;                      // longer reads take the unstaged byte path
constexpr int RUN_MAX = 2 * COPY_MAXQ + 1;
constexpr int RUN_ROW = (RUN_MAX + 3 + 8) & ~3;                // word-aligned rows, slack for the 5-word window
struct CopySmem {
  uint8_t fwd[WARPS][RUN_ROW];                                 // SEQ \t QUAL as given
  uint8_t rev[WARPS][RUN_ROW];                                 // reverse complement \t reversed QUAL
  uint8_t comp[256];                                           // reverse_complement's character map (fasta.cpp:26-61)
};
constexpr int COPY_WARPS = 8;
static_assert(COPY_WARPS == WARPS, "CopySmem is sized by WARPS");

__device__ __forceinline__ void copy_run(char *__restrict__ dst, const uint8_t *__restrict__ src, int len, int lane) {
  const int head = (int)((16u - (unsigned)((uintptr_t)dst & 15u)) & 15u);
  const int hl = head < len ? head : len;
  const int nb = (len - hl) >> 4, tail = (len - hl) & 15;
  {                                                            // edges: one byte store
    const int t = lane & 15;
    const int i = lane < 16 ? t : hl + 16 * nb + t;
    if (lane < 16 ? t < hl : t < tail) dst[i] = (char)src[i];
  }
  const unsigned sh = (unsigned)(hl & 3) * 8u;
  for (int k = lane; k < nb; k += 32) {
    const int s0 = hl + 16 * k;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(src + (s0 & ~3));   // rows are word-aligned
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
    uint4 v;
    v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
    v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
    *reinterpret_cast<uint4 *>(dst + s0) = v;
  }
}

#ifndef SMASH_COPY_MINBLK
#define SMASH_COPY_MINBLK 6
#endif
__global__ void __launch_bounds__(THREADS, SMASH_COPY_MINBLK)
k_emit_copy(BatchDev b, WorkDev w) {
  extern __shared__ __align__(16) uint8_t copy_smem_raw[];
  CopySmem &sm = *reinterpret_cast<CopySmem *>(copy_smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  sm.comp[threadIdx.x & 255] = comp_char((uint8_t)(threadIdx.x & 255));
  __syncthreads();
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  // a warp takes 32 consecutive reads: lane l fetches the placement data of read l (coalesced, one round
  // of latency per 32 reads), then the warp walks the reads with shuffles
  for (uint64_t blk = (uint64_t)blockIdx.x * WARPS + warp; blk * 32 < b.n_reads; blk += warps_total) {
    const uint64_t rd = blk * 32 + (uint64_t)lane;
    int m_nrec = 0; uint64_t m_fbase = 0; int64_t m_so = 0, m_no = 0; int m_q = 0, m_nl = 0; int m_unm = 0;
    if (rd < b.n_reads) {
      m_nrec = (int)w.nrec[rd]; m_fbase = w.rec_base[rd];
      m_so = b.seq_off[rd]; m_q = (int)(b.seq_off[rd + 1] - m_so);
      m_no = b.name_off[rd]; m_nl = (int)(b.name_off[rd + 1] - m_no);
      m_unm = w.sums[rd].unmapped ? 1 : 0;
    }
  for (int ri = 0; ri < 32; ++ri) {
    const int n_rec = __shfl_sync(0xffffffffu, m_nrec, ri);
    if (!n_rec) continue;
    const uint64_t read = blk * 32 + (uint64_t)ri;
    const uint64_t fbase = __shfl_sync(0xffffffffu, m_fbase, ri);
    const int64_t so = __shfl_sync(0xffffffffu, m_so, ri), no = __shfl_sync(0xffffffffu, m_no, ri);
    const int q = __shfl_sync(0xffffffffu, m_q, ri), name_len = __shfl_sync(0xffffffffu, m_nl, ri);
    const uint8_t *__restrict__ name = b.names + no;
    const bool unmapped = __shfl_sync(0xffffffffu, m_unm, ri) != 0;
    const Rec *recs = w.rec_slots + slot_base(w, read);
    const bool staged = q <= COPY_MAXQ;
    const uint8_t nm = lane < name_len ? name[lane] : 0;
    // lanes = first 32 records: their placement, requested together with the read's bytes
    uint64_t my_off = 0; uint32_t my_meta = 0, my_tail = 0;
    if (lane < n_rec) {
      const Rec rr = recs[lane];
      my_off = w.rec_off[fbase + (uint64_t)lane];
      my_meta = (uint32_t)rr.seq_off | (rr.rc && !unmapped ? 0x80000000u : 0u);
      if (b.opt) my_tail = w.rec_bytes[fbase + (uint64_t)lane] - 1u - rr.lr_len;
    }
    if (staged) {
      // word-wise fetch of SEQ and QUAL (same misalignment: they share seq_off), all words of a
      // 256-byte pass in flight together; bytes are scattered into the forward and reversed runs
      uint8_t *F = sm.fwd[warp], *R = sm.rev[warp];
      const int mis = (int)(so & 3);
      const uint32_t *sw = reinterpret_cast<const uint32_t *>(b.seq + (so - mis));
      const uint32_t *qw = reinterpret_cast<const uint32_t *>(b.qual + (so - mis));
      const int nwords = (mis + q + 3) >> 2;
      for (int w0 = 0; w0 < nwords; w0 += 64) {
        const int i0 = w0 + lane, i1 = w0 + 32 + lane;
        uint32_t s0 = 0, q0 = 0, s1 = 0, q1 = 0;
        if (i0 < nwords) { s0 = __ldg(sw + i0); q0 = __ldg(qw + i0); }
        if (i1 < nwords) { s1 = __ldg(sw + i1); q1 = __ldg(qw + i1); }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = h ? i1 : i0;
          const uint32_t sv4 = h ? s1 : s0, qv4 = h ? q1 : q0;
          if (i < nwords) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int j = 4 * i - mis + t;
              if (j >= 0 && j < q) {
                const uint8_t sv = (uint8_t)(sv4 >> (8 * t)), qv = (uint8_t)(qv4 >> (8 * t));
                F[j] = sv; F[q + 1 + j] = qv;
                R[q - 1 - j] = sm.comp[sv]; R[2 * q - j] = qv;
              }
            }
          }
        }
      }
      if (lane == 0) { F[q] = '\t'; R[q] = '\t'; }
    }
    __syncwarp();
    for (int r0 = 0; r0 < n_rec; r0 += 32) {
      if (r0) {                                              // more than 32 records: next block's placement
        my_off = 0; my_meta = 0; my_tail = 0;
        if (r0 + lane < n_rec) {
          const Rec rr = recs[r0 + lane];
          my_off = w.rec_off[fbase + (uint64_t)(r0 + lane)];
          my_meta = (uint32_t)rr.seq_off | (rr.rc && !unmapped ? 0x80000000u : 0u);
          if (b.opt) my_tail = w.rec_bytes[fbase + (uint64_t)(r0 + lane)] - 1u - rr.lr_len;
        }
      }
      const int nr = n_rec - r0 < 32 ? n_rec - r0 : 32;
      for (int r = 0; r < nr; ++r) {
        const uint64_t off = __shfl_sync(0xffffffffu, my_off, r);
        const uint32_t meta = __shfl_sync(0xffffffffu, my_meta, r);
        char *out = w.sam + off;
        const bool rc = (meta >> 31) != 0;
        if (lane < name_len) out[lane] = (char)nm;
        for (int i = lane + 32; i < name_len; i += 32) out[i] = (char)name[i];
        char *o2 = out + (meta & 0x7fffffffu);
        if (staged) {
          copy_run(o2, rc ? sm.rev[warp] : sm.fwd[warp], 2 * q + 1, lane);
        } else {
          const uint8_t *__restrict__ seq = b.seq + so, *__restrict__ qual = b.qual + so;
          for (int j = lane; j < q; j += 32) {
            const int src = rc ? q - 1 - j : j;
            o2[j] = (char)(rc ? sm.comp[seq[src]] : seq[src]); o2[q + 1 + j] = (char)qual[src];
          }
          if (lane == 0) o2[q] = '\t';
        }
        if (b.opt) {
          const uint32_t tail = __shfl_sync(0xffffffffu, my_tail, r);
          const int64_t oo = b.opt_off[read];
          const int opt_len = (int)(b.opt_off[read + 1] - oo);
          char *o3 = out + (tail - (uint32_t)opt_len);
          for (int i = lane; i < opt_len; i += 32) o3[i] = (char)b.opt[oo + i];
        }
      }
    }
    __syncwarp();
  }
  }
}

int launch_emit_text(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st, uint64_t n_records) {
  if (!b.n_reads || !n_records) return 0;
  const uint64_t need = (n_records + 127) / 128, cap = (uint64_t)sm_count() * SMASH_TEXT_CTAS;
  k_emit_text<<<(unsigned)(need < cap ? need : cap), 128, 0, st>>>(ix, b, w, p, n_records);
  return 1;
}
int launch_emit_compact(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st, uint64_t n_records) {
  if (!b.n_reads || !n_records) return 0;
  const uint64_t need = (n_records + 127) / 128, cap = (uint64_t)sm_count() * 16;
  k_emit_compact<<<(unsigned)(need < cap ? need : cap), 128, 0, st>>>(ix, b, w, p, n_records);
  return 1;
}
int launch_emit_copy(const BatchDev &b, const WorkDev &w, cudaStream_t st, uint64_t n_records) {
  if (!b.n_reads || !n_records) return 0;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_emit_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CopySmem)); attr_set = true; }
  k_emit_copy<<<grid_for_warps((b.n_reads + 31) / 32, SMASH_COPY_MINBLK), THREADS, sizeof(CopySmem), st>>>(b, w);
  return 1;
}

// ------------------------------------------------------------------ K5: record_sort
//
// A reference worker sorts the lines of a chunk before it writes the file (OutputSorter::flush, query.cpp:448-468) with
// MemSam::operator< (memsam.h:136-158): absolute position = MemSam::chromosomes[RNAME] + POS as PRINTED (an unmapped
// placeholder carries its mate's RNAME/POS, "*" sorts after every chromosome, query.cpp:546-552), then the name
// (std::string <, bytes as unsigned), then flag & (first | second | reversed); two lines equal in all three make the
// reference throw "flags equal".  The order is decided before any text exists: keys from the records, a merge sort of
// the flat record indices, and rec_off re-derived, so k_emit_* write every line straight into its sorted place.
__global__ void __launch_bounds__(128)
k_sort_keys(DevIndex ix, BatchDev b, WorkDev w, uint64_t n_records) {
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_records; f += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t read = w.rec_read[f];
    const int r = (int)(f - w.rec_base[read]);
    const ReadSum me = w.sums[read];
    uint16_t flag; MateView mv;
    read_mate(b, w, read, &flag, &mv);
    const int n_fwd = ix.rcref ? ix.n_descr / 2 : ix.n_descr;
    uint64_t abs;
    if (me.unmapped) {
      abs = mv.has ? ix.chrom_abs64[ix.rcref ? mv.si >> 1 : mv.si] + (uint64_t)(mv.pos + 1) : ix.chrom_abs64[n_fwd];
    } else {
      const Rec rr = (w.rec_slots + slot_base(w, read))[r];
      abs = ix.chrom_abs64[ix.rcref ? rr.si >> 1 : rr.si] + (uint64_t)(rr.pos + 1);
      flag = (uint16_t)(flag | (rr.rc ? 16 : 0));
    }
    w.sort_abs[f] = abs;
    w.sort_flag[f] = (uint8_t)(flag & (64 | 128 | 16));
    w.sort_perm[f] = (uint32_t)f;
  }
}
struct RecLess {
  const uint64_t *abs; const uint8_t *fl; const uint32_t *rec_read; const uint8_t *names; const int64_t *name_off;
  __host__ __device__ int cmp_name(uint32_t x, uint32_t y) const {
    const uint32_t rx = rec_read[x], ry = rec_read[y];
    if (rx == ry) return 0;
    const uint8_t *a = names + name_off[rx], *c = names + name_off[ry];
    const int64_t la = name_off[rx + 1] - name_off[rx], lc = name_off[ry + 1] - name_off[ry];
    const int64_t m = la < lc ? la : lc;
    for (int64_t i = 0; i < m; ++i) if (a[i] != c[i]) return a[i] < c[i] ? -1 : 1;
    return la < lc ? -1 : la > lc ? 1 : 0;
  }
  __host__ __device__ bool operator()(uint32_t x, uint32_t y) const {
    if (abs[x] != abs[y]) return abs[x] < abs[y];
    const int c = cmp_name(x, y);
    if (c) return c < 0;
    if (fl[x] != fl[y]) return fl[x] < fl[y];
    return x < y;
  }
};
__global__ void k_sort_scatter(WorkDev w, RecLess less, uint64_t n) {
  unsigned dup = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t f = w.sort_perm[i];
    w.sort_bytes[i] = w.rec_bytes[f];
    if (i) {                                                   // "flags equal" (memsam.h:143-150)
      const uint32_t g = w.sort_perm[i - 1];
      if (less.abs[f] == less.abs[g] && less.fl[f] == less.fl[g] && less.cmp_name(f, g) == 0) ++dup;
    }
  }
  if (dup) atomicAdd(&w.flags[FLAG_SORTDUP], dup);
}
__global__ void k_sort_gather(WorkDev w, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    w.rec_off[w.sort_perm[i]] = w.sort_off[i];
}
int launch_record_sort(const DevIndex &ix, const BatchDev &b, const WorkDev &w, uint64_t n, cudaStream_t st, size_t *tmp_bytes_needed) {
  const RecLess less{w.sort_abs, w.sort_flag, w.rec_read, b.names, b.name_off};
  if (tmp_bytes_needed) {
    *tmp_bytes_needed = 0;
#if !defined(SMASH_CUDA_SHIM)
    cub::DeviceMergeSort::SortKeys(nullptr, *tmp_bytes_needed, (uint32_t *)nullptr, (int64_t)n, less, st);
#endif
    return 0;
  }
  if (!n) return 0;
  const uint64_t need = (n + 127) / 128, cap = (uint64_t)sm_count() * 16;
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  k_sort_keys<<<grid, 128, 0, st>>>(ix, b, w, n);
#if !defined(SMASH_CUDA_SHIM)
  size_t tb = w.sort_tmp_bytes;
  cub::DeviceMergeSort::SortKeys(w.sort_tmp, tb, w.sort_perm, (int64_t)n, less, st);
#else
  cudaStreamSynchronize(st);
  std::sort(w.sort_perm, w.sort_perm + n, less);             // host emulation (tests/emul)
#endif
  k_sort_scatter<<<grid, 128, 0, st>>>(w, less, n);
  const int ns = exclusive_scan_u32(w.sort_bytes, n, w.blk_sums2, w.sort_off, st);
  k_sort_gather<<<grid, 128, 0, st>>>(w, n);
  return 4 + ns;
}

// ------------------------------------------------------------------ matches -> CSR for the host

__global__ void k_match_copy(BatchDev b, WorkDev w, const uint64_t *__restrict__ off, uint64_t *__restrict__ triples) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t read = (uint64_t)blockIdx.x * WARPS + warp; read < b.n_reads; read += warps_total) {
    const uint32_t n = w.match_cnt[read];
    const Match *src = w.match_slots + slot_base(w, read);
    uint64_t *dst = triples + 3 * off[read];
    for (uint32_t i = lane; i < n && i < (uint32_t)slot_cap(w, read); i += 32) {
      dst[3 * i] = src[i].ref; dst[3 * i + 1] = src[i].qpos; dst[3 * i + 2] = src[i].len;
    }
  }
}
int launch_match_csr(const BatchDev &b, const WorkDev &w, int64_t *off, uint64_t *triples, uint64_t *scratch, cudaStream_t st) {
  if (!b.n_reads) return 0;
  int n = exclusive_scan_u32(w.match_cnt, b.n_reads, scratch, (uint64_t *)off, st);
  k_match_copy<<<grid_for_warps(b.n_reads, 8), THREADS, 0, st>>>(b, w, (const uint64_t *)off, triples);
  return n + 1;
}

// ------------------------------------------------------------------ map.bin on the GPU

// longSA::show, bin=true (longSA.cpp:612-690): two bytes per forward base.  The reference fills an
// N-entry vector in SA order and then edits single entries; every entry is visited once, so the
// result is a pure function of (SA, ISA, LCP) that is evaluated here per base, with no scratch.
__device__ __forceinline__ uint64_t min_unique_len(const DevIndex &ix, uint64_t sa_index) {
  const uint64_t a = lcp_at(ix, sa_index);
  const uint64_t b = sa_index + 1 < ix.N ? lcp_at(ix, sa_index + 1) : 0;
  return (a > b ? a : b) + 1;
}
__global__ void k_mappability(DevIndex ix, int chrom, uint64_t out_base, uint8_t *__restrict__ body) {
  const uint64_t start = ix.startpos[chrom], size = ix.sizes[chrom];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = min_unique_len(ix, isa_at(ix, start + i));
    uint64_t l = min_unique_len(ix, isa_at(ix, start + 2 * size - i));
    if (r + i >= size) r = 0;
    if (l >= i) l = 0;
    body[out_base + 2 * i] = (uint8_t)(l < 255 ? l : 255);
    body[out_base + 2 * i + 1] = (uint8_t)(r < 255 ? r : 255);
  }
}
int launch_mappability(const DevIndex &ix, uint64_t *, uint8_t *body, cudaStream_t st) {
  // host-side loop over forward chromosomes; startpos/sizes are read on the device
  int launches = 0;
  uint64_t base = 0;
  uint64_t *h_sizes = new uint64_t[ix.n_descr];
  cudaMemcpyAsync(h_sizes, ix.sizes, sizeof(uint64_t) * ix.n_descr, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  for (int c = 0; c < ix.n_descr; c += 2) {
    k_mappability<<<sm_count() * 8, 256, 0, st>>>(ix, c, base, body);
    base += 2 * h_sizes[c];
    ++launches;
  }
  delete[] h_sizes;
  return launches;
}

__global__ void k_add_one(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i] + 1u;
}
int launch_slot_offsets(const uint32_t *match_cnt, uint64_t n_reads, uint32_t *tmp, uint64_t *blk, uint64_t *slot_off, cudaStream_t st) {
  if (!n_reads) return 0;
  k_add_one<<<sm_count() * 4, 256, 0, st>>>(match_cnt, n_reads, tmp);
  return 1 + exclusive_scan_u32(tmp, n_reads, blk, slot_off, st);
}

}  // namespace smash

namespace smash {
int exclusive_scan_u32_public(const uint32_t *in, uint64_t n, uint64_t *blk, uint64_t *out, cudaStream_t st) {
  return exclusive_scan_u32(in, n, blk, out, st, nullptr);
}
}  // namespace smash
