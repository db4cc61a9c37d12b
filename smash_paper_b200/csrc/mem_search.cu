// mem_search.cu -- K2: longSA::MEM / findMEM (-maxmatch), longSA.cpp:395-490.
//
// MEM mode is NOT free in its algorithm the way MAM is: the output
// depends on the control flow -- the search starts at query offset 1 (`prefix = 1`, longSA.cpp:398),
// the return value of `suffixlink(&xmi)` is ignored (longSA.cpp:424) so a failed `expand_link`
// (threshold 2*depth*logN, longSA.h:159) leaves a shrunken interval behind, and the emission order
// (SA order inside the max-match interval, then left/right LCP unwinding per level) decides how
// libstdc++'s std::sort breaks ties later.  So this kernel restates that control flow faithfully:
// ONE THREAD PER READ walks the query with the ISA/LCP suffix-link simulation exactly as the
// reference does, first in a counting pass, then (after a prefix sum) in a writing pass that emits
// the matches in the reference's order into a CSR.  Divergence is inherent; MEM mode is the
// secondary mode of the pipeline (production runs MAM, smash_mapping.sh:19).
#include "kernels.cuh"

namespace smash {

struct Ivl { uint64_t depth, lo, hi; };     // interval_t (longSA.h:64-75)

struct QueryView {                           // lower-cased on the fly (NewQuery::extend, query.cpp:125-144)
  const uint8_t *seq; int q; int nuc;
  __device__ __forceinline__ uint8_t operator[](uint64_t i) const { return query_char(seq[i], nuc); }
};

// top_down_faster (longSA.cpp:322-380): first/last suffix of [lo,hi] whose character at `off` is c
__device__ __forceinline__ bool narrow(const DevIndex &ix, uint8_t c, uint64_t off, uint64_t *lo, uint64_t *hi) {
  const uint8_t *T = ix.text;
  const uint64_t a = *lo, b = *hi;
  if (c < T[sa_at(ix, a) + off] || c > T[sa_at(ix, b) + off]) return false;
  uint64_t l = a, r = b + 1;
  while (l < r) { const uint64_t m = l + ((r - l) >> 1); if (T[sa_at(ix, m) + off] < c) l = m + 1; else r = m; }
  const uint64_t first = l;
  if (first > b || T[sa_at(ix, first) + off] != c) return false;
  l = first; r = b + 1;
  while (l < r) { const uint64_t m = l + ((r - l) >> 1); if (T[sa_at(ix, m) + off] <= c) l = m + 1; else r = m; }
  *lo = first; *hi = l - 1;
  return true;
}
// traverse (longSA.cpp:297-316)
__device__ __forceinline__ void traverse(const DevIndex &ix, const QueryView &P, uint64_t prefix, Ivl *cur, uint64_t stop_len) {
  if (cur->depth >= stop_len) return;
  while (prefix + cur->depth < (uint64_t)P.q) {
    uint64_t lo = cur->lo, hi = cur->hi;
    if (!narrow(ix, P[prefix + cur->depth], cur->depth, &lo, &hi)) return;
    cur->depth += 1; cur->lo = lo; cur->hi = hi;
    if (cur->depth == stop_len) return;
  }
}
// expand_link (longSA.h:158-174)
__device__ __forceinline__ bool expand_link(const DevIndex &ix, Ivl *v) {
  const uint64_t thresh = 2 * v->depth * ix.logN;
  uint64_t steps = 0, lo = v->lo, hi = v->hi;
  while (lcp_at(ix, lo) >= v->depth) { if (++steps >= thresh) return false; --lo; }
  while (hi < ix.N - 1 && lcp_at(ix, hi + 1) >= v->depth) { if (++steps >= thresh) return false; ++hi; }
  v->lo = lo; v->hi = hi;
  return true;
}
// suffixlink (longSA.cpp:383-392)
__device__ __forceinline__ bool suffixlink(const DevIndex &ix, Ivl *v) {
  if (v->depth <= 1) { v->depth = 0; return false; }
  v->depth -= 1;
  v->lo = isa_at(ix, sa_at(ix, v->lo) + 1);
  v->hi = isa_at(ix, sa_at(ix, v->hi) + 1);
  return expand_link(ix, v);
}

// WRITE: matches go to their CSR place.  !WRITE (the counting pass): the first `cap` matches of the read are kept in a
// small per-read stage, so that a read with at most MEM_STAGE matches -- nearly all of them -- is not searched a second
// time: the writing pass copies its staged matches to the CSR place and redoes only the reads that overflowed the stage.
template <bool WRITE> struct Sink {
  Match *out; uint64_t n; uint64_t cap;
  __device__ __forceinline__ void emit(uint64_t ref, uint64_t qpos, uint64_t len) {
    if (WRITE || n < cap) { out[n].ref = ref; out[n].qpos = (uint32_t)qpos; out[n].len = (uint32_t)len; }
    ++n;
  }
};
// find_Lmaximal (longSA.cpp:438-457)
template <bool WRITE>
__device__ __forceinline__ void left_maximal_emit(const DevIndex &ix, const QueryView &P, uint64_t min_len, uint64_t prefix,
                                                  uint64_t r, uint64_t len, Sink<WRITE> *s) {
  if (prefix == 0 || r == 0 || P[prefix - 1] != ix.text[r - 1])
    if (len >= min_len) s->emit(r, prefix, len);
}
// collectMEMs (longSA.cpp:461-490)
template <bool WRITE>
__device__ void collect_mems(const DevIndex &ix, const QueryView &P, uint64_t min_len, uint64_t prefix, Ivl mli, Ivl xmi, Sink<WRITE> *s) {
  for (uint64_t i = xmi.lo; i <= xmi.hi; ++i) left_maximal_emit(ix, P, min_len, prefix, sa_at(ix, i), xmi.depth, s);
  if (mli.lo == xmi.lo && mli.hi == xmi.hi) return;
  while (xmi.depth >= mli.depth) {
    if (xmi.hi + 1 < ix.N) {
      const uint64_t a = lcp_at(ix, xmi.lo), b = lcp_at(ix, xmi.hi + 1);
      xmi.depth = a > b ? a : b;
    } else {
      xmi.depth = lcp_at(ix, xmi.lo);
    }
    if (xmi.depth >= mli.depth) {
      while (lcp_at(ix, xmi.lo) >= xmi.depth) {
        --xmi.lo;
        left_maximal_emit(ix, P, min_len, prefix, sa_at(ix, xmi.lo), xmi.depth, s);
      }
      while (xmi.hi + 1 < ix.N && lcp_at(ix, xmi.hi + 1) >= xmi.depth) {
        ++xmi.hi;
        left_maximal_emit(ix, P, min_len, prefix, sa_at(ix, xmi.hi), xmi.depth, s);
      }
    }
  }
}
// findMEM (longSA.cpp:395-435)
template <bool WRITE>
__device__ void find_mem(const DevIndex &ix, const QueryView &P, uint64_t min_len, Sink<WRITE> *s) {
  if (min_len < 1) return;                                   // longSA::MEM, longSA.cpp:587-590
  const uint64_t last = ix.N - 1, q = (uint64_t)P.q;
  uint64_t prefix = 1;
  Ivl mli = {0, 0, last}, xmi = {0, 0, last};
  while (prefix <= q) {
    traverse(ix, P, prefix, &mli, min_len);
    if (mli.depth > xmi.depth) xmi = mli;
    if (mli.depth <= 1) {
      mli.depth = 0; mli.lo = 0; mli.hi = last; xmi = mli;
      ++prefix;
      continue;
    }
    if (mli.depth >= min_len) {
      traverse(ix, P, prefix, &xmi, q);
      collect_mems(ix, P, min_len, prefix, mli, xmi, s);
      ++prefix;
      if (!suffixlink(ix, &mli)) { mli.depth = 0; mli.lo = 0; mli.hi = last; xmi = mli; continue; }
      (void)suffixlink(ix, &xmi);
    } else {
      ++prefix;
      if (!suffixlink(ix, &mli)) { mli.depth = 0; mli.lo = 0; mli.hi = last; xmi = mli; continue; }
      xmi = mli;
    }
  }
}

// Launch geometry, measured on config 1 (1 M reads, r02t/r02u A/B): one thread per read with the grid uncapped (the block
// scheduler balances the very uneven reads; a capped grid-stride grid leaves warps waiting for their slowest thread
// several reads in a row) and 32 registers (16 CTAs of 128 threads per SM; the walk is latency-bound, the few spills cost
// less than the extra warps gain): 81.9 -> 62.2 ms.
#ifndef SMASH_MEM_GRID_CAP
#define SMASH_MEM_GRID_CAP 0x7fffffff
#endif
constexpr uint64_t MEM_GRID_CAP = SMASH_MEM_GRID_CAP;
#ifndef SMASH_MEM_MINBLK
#define SMASH_MEM_MINBLK 16
#endif
template <bool WRITE>
__global__ void __launch_bounds__(128, SMASH_MEM_MINBLK)
k_mem_search(DevIndex ix, BatchDev b, SearchParams sp, uint32_t min_len_raw, uint32_t *__restrict__ cnt,
             const uint64_t *__restrict__ off, Match *__restrict__ matches, Match *__restrict__ stage) {
  for (uint64_t read = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; read < b.n_reads; read += (uint64_t)gridDim.x * blockDim.x) {
    const int64_t so = b.seq_off[read];
    QueryView P{b.seq + so, (int)(b.seq_off[read + 1] - so), sp.nucleotides_only};
    if (WRITE) {
      const uint64_t o = off[read], n = cnt[read];
      if (stage && n <= (uint64_t)MEM_STAGE) {                       // searched already: the counting pass kept its matches
        for (uint64_t i = 0; i < n; ++i) matches[o + i] = stage[read * MEM_STAGE + i];
        continue;
      }
      Sink<WRITE> s{matches + o, 0, 0};
      find_mem<WRITE>(ix, P, (uint64_t)min_len_raw, &s);
    } else {
      Sink<WRITE> s{stage ? stage + read * MEM_STAGE : nullptr, 0, stage ? (uint64_t)MEM_STAGE : 0};
      find_mem<WRITE>(ix, P, (uint64_t)min_len_raw, &s);
      cnt[read] = (uint32_t)(s.n > 0xfffffffeull ? 0xfffffffeull : s.n);
    }
  }
}

int launch_mem_count(const DevIndex &ix, const BatchDev &b, const SearchParams &p, uint32_t min_len_raw, uint32_t *cnt, Match *stage,
                     cudaStream_t st) {
  if (!b.n_reads) return 0;
  const uint64_t need = (b.n_reads + 127) / 128;
  k_mem_search<false><<<(unsigned)(need < MEM_GRID_CAP ? need : MEM_GRID_CAP), 128, 0, st>>>(ix, b, p, min_len_raw, cnt, nullptr, nullptr, stage);
  return 1;
}
int launch_mem_write(const DevIndex &ix, const BatchDev &b, const SearchParams &p, uint32_t min_len_raw, const uint64_t *off,
                     Match *matches, const Match *stage, const uint32_t *cnt, cudaStream_t st) {
  if (!b.n_reads) return 0;
  const uint64_t need = (b.n_reads + 127) / 128;
  k_mem_search<true><<<(unsigned)(need < MEM_GRID_CAP ? need : MEM_GRID_CAP), 128, 0, st>>>(ix, b, p, min_len_raw, const_cast<uint32_t *>(cnt), off,
                                                                                   matches, const_cast<Match *>(stage));
  return 1;
}

}  // namespace smash
