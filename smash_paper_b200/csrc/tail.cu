// tail.cu -- K7 smash_filter (smashMEM.py:84-92, 193-228 with argv "0 0 10000 4") and
// K8 varbin_hist (varbin.py:34-93) on the GPU.
//
// Per batch (tail_accumulate): one thread per read pair applies the excess-mappability filter
// and the read-2-near-read-1 window, and appends the pair's kept hits (tid,pos in HI order, r1
// then r2) plus a 128-bit fingerprint of the dupe key to HBM-resident arrays.
// At the end (tail_finish): first-wins duplicate removal through an open-addressing table keyed
// by the fingerprint with an exact comparison of the hit lists, ordered compaction of the
// surviving hits into the positions list (prefix scan => the reference's output order), the
// chromosome filters, varbin's adjacent-duplicate rule and the bin histogram
// (shared-memory privatised when the bins fit in one SM's shared memory).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "tail.cuh"
#if !defined(SMASH_CUDA_SHIM)
#include <cub/device/device_merge_sort.cuh>
#else
#include <algorithm>
#endif
#include <chrono>
#include <stdlib.h>
static const bool t_dbg = getenv("SMASH_DEBUG_TIMING") != nullptr;
static const bool t_dbg_sync = getenv("SMASH_DEBUG_TIMING") && atoi(getenv("SMASH_DEBUG_TIMING")) >= 2;
static double t_now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TDBG(label) do { if (t_dbg) { if (t_dbg_sync) cudaStreamSynchronize(st); fprintf(stderr, "[smash-dbg]    finish:%-20s %8.3f ms\n", label, t_now() - tq); tq = t_now(); } } while (0)

namespace smash {

static thread_local char t_err[512] = "";
const char *tail_error() { return t_err; }
static int tfail(int code, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(t_err, sizeof t_err, fmt, ap); va_end(ap);
  return code;
}
#define TCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return tfail(SMASH_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

template <class T> int DGrow<T>::reserve(size_t n, size_t used, cudaStream_t st) {
  if (n <= cap) return 0;
  size_t want = cap * 2 > n ? cap * 2 : n + n / 8 + 1024;
  T *np = nullptr;
  cudaError_t e = cudaMalloc((void **)&np, want * sizeof(T));
  if (e != cudaSuccess) return tfail(SMASH_ERR_NOMEM, "cudaMalloc(%zu): %s", want * sizeof(T), cudaGetErrorString(e));
  if (p && used) {
    TCU(cudaMemcpyAsync(np, p, used * sizeof(T), cudaMemcpyDeviceToDevice, st));
    TCU(cudaStreamSynchronize(st));
  }
  if (p) cudaFree(p);
  p = np; cap = want;
  return 0;
}
template struct DGrow<uint32_t>;
template struct DGrow<uint8_t>;
template struct DGrow<uint64_t>;

// device-side running totals (t->d_nhits): kept hits, bytes in name_blob, pairs out of name order
enum { CNT_HITS = 0, CNT_NAME_BYTES = 1, CNT_ORDER_VIOLATIONS = 2, N_TAIL_COUNTERS = 4 };

void tail_init(TailState *) {}
void tail_reset(TailState *t) {
  t->n_pairs = 0; t->n_hits_bound = 0; t->n_name_bound = 0; t->n_positions = 0; t->order_violations = 0;
  if (t->d_nhits) cudaMemset(t->d_nhits, 0, 8 * N_TAIL_COUNTERS);
}
void tail_release(TailState *t) {
  t->pair_nhits.release(); t->pair_fp.release(); t->pair_hit_off.release(); t->hits.release();
  t->name_blob.release(); t->pair_name_off.release();
  for (auto &b : t->scr) b.release();
  t->exp_keys.release();
  void *d[] = {t->bin_starts, t->chrom_off, t->batch_cnt, t->batch_off, t->blk, t->counts, t->pos_chrom, t->pos_pos, t->d_nhits, t->bin_lut};
  for (void *p : d) if (p) cudaFree(p);
  if (t->h_pos_chrom) cudaFreeHost(t->h_pos_chrom);
  if (t->h_pos_pos) cudaFreeHost(t->h_pos_pos);
  if (t->last_ev) cudaEventDestroy(t->last_ev);
  *t = TailState();
}

int tail_configure(TailState *t, const int64_t *bin_starts, uint64_t n_bins, const int64_t *chrom_off,
                   uint64_t n_chrom, int64_t hit_window, int32_t min_excess) {
  if (t->bin_starts) cudaFree(t->bin_starts);
  if (t->chrom_off) cudaFree(t->chrom_off);
  if (t->counts) cudaFree(t->counts);
  t->bin_starts = nullptr; t->chrom_off = nullptr; t->counts = nullptr;
  TCU(cudaMalloc((void **)&t->bin_starts, 8 * n_bins));
  TCU(cudaMalloc((void **)&t->chrom_off, 8 * (n_chrom + 1)));
  TCU(cudaMalloc((void **)&t->counts, 8 * n_bins));
  TCU(cudaMemcpy(t->bin_starts, bin_starts, 8 * n_bins, cudaMemcpyHostToDevice));
  // granule table for bin_of: lut[g] = number of starts <= g << shift, at most 2^20 granules; only for sorted,
  // non-negative starts (anything else keeps the plain bisect, whose answer on unsorted input is what python's is)
  if (t->bin_lut) { cudaFree(t->bin_lut); t->bin_lut = nullptr; }
  t->lut_n = 0; t->lut_shift = 0;
  bool sorted = n_bins > 0 && n_bins < 0xffffffffull && bin_starts[0] >= 0;
  for (uint64_t i = 1; sorted && i < n_bins; ++i) sorted = bin_starts[i - 1] <= bin_starts[i];
  if (sorted) {
    const uint64_t top = (uint64_t)bin_starts[n_bins - 1];
    int shift = 0;
    while ((top >> shift) + 2 > (1ull << 20)) ++shift;
    const uint64_t n = (top >> shift) + 2;
    std::vector<uint32_t> lut(n);
    uint64_t c = 0;
    for (uint64_t g = 0; g < n; ++g) {
      const uint64_t v = g << shift;
      while (c < n_bins && (uint64_t)bin_starts[c] <= v) ++c;
      lut[g] = (uint32_t)c;
    }
    TCU(cudaMalloc((void **)&t->bin_lut, 4 * n));
    TCU(cudaMemcpy(t->bin_lut, lut.data(), 4 * n, cudaMemcpyHostToDevice));
    t->lut_n = n; t->lut_shift = shift;
  }
  TCU(cudaMemcpy(t->chrom_off, chrom_off, 8 * n_chrom, cudaMemcpyHostToDevice));
  t->n_bins = n_bins; t->n_chrom = n_chrom; t->hit_window = hit_window; t->min_excess = min_excess;
  t->configured = true;
  tail_reset(t);
  return 0;
}

// ------------------------------------------------------------------ per batch

struct PairParams { int64_t hit_window; int32_t min_excess; };

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
  return x;
}
// smashMEM.py:84-92: mapped and qlen - max(L0,R0) >= minExcessMappability
__device__ __forceinline__ bool rec_passes(const Rec &r, int32_t min_excess) {
  const int mm = r.L0 > r.R0 ? r.L0 : r.R0;
  return (int)r.qlen - mm >= min_excess;
}
// Visit the kept hits of pair (2p, 2p+1) in output order; f(tid, pos).
template <class F>
__device__ __forceinline__ void for_kept_hits(const BatchDev &b, const WorkDev &w, uint64_t p, const PairParams &pp, F f) {
  const uint64_t ra = 2 * p, rb = 2 * p + 1;
  const ReadSum sa = w.sums[ra];
  const Rec *reca = w.rec_slots + slot_base(w, ra);
  const int na = sa.unmapped ? 0 : sa.n_rec;
  for (int i = 0; i < na; ++i)
    if (rec_passes(reca[i], pp.min_excess)) f((uint32_t)(reca[i].si >> 1), reca[i].pos);
  if (rb >= b.n_reads) return;
  const ReadSum sb = w.sums[rb];
  const Rec *recb = w.rec_slots + slot_base(w, rb);
  const int nb = sb.unmapped ? 0 : sb.n_rec;
  for (int j = 0; j < nb; ++j) {
    if (!rec_passes(recb[j], pp.min_excess)) continue;
    bool near = false;                                       // smashMEM.py:193-208
    for (int i = 0; i < na && !near; ++i) {
      if (!rec_passes(reca[i], pp.min_excess)) continue;
      if (reca[i].si != recb[j].si) continue;
      int64_t d = reca[i].pos - recb[j].pos; if (d < 0) d = -d;
      near = d < pp.hit_window;
    }
    if (!near) f((uint32_t)(recb[j].si >> 1), recb[j].pos);
  }
}

__global__ void k_pair_count(BatchDev b, WorkDev w, PairParams pp, uint64_t n_pairs, uint32_t *__restrict__ cnt) {
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t n = 0;
    for_kept_hits(b, w, p, pp, [&](uint32_t, int64_t) { ++n; });
    cnt[p] = n;
  }
}
__global__ void k_pair_write(BatchDev b, WorkDev w, PairParams pp, uint64_t n_pairs, const uint64_t *__restrict__ off,
                             uint64_t pair_base, const uint64_t *__restrict__ hit_base_p, uint32_t *__restrict__ pair_nhits,
                             uint64_t *__restrict__ pair_fp, uint64_t *__restrict__ pair_hit_off, uint64_t *__restrict__ hits) {
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t o = *hit_base_p + off[p];
    const uint64_t o0 = o;
    uint64_t h1 = 0x243f6a8885a308d3ULL, h2 = 0x13198a2e03707344ULL;
    for_kept_hits(b, w, p, pp, [&](uint32_t tid, int64_t pos) {
      const uint64_t v = ((uint64_t)tid << 40) | (uint64_t)pos;
      hits[o++] = v;
      h1 = mix64(h1 ^ v) + 0x9e3779b97f4a7c15ULL;
      h2 = mix64((h2 + v) * 0xff51afd7ed558ccdULL) ^ (h2 >> 29);
    });
    pair_nhits[pair_base + p] = (uint32_t)(o - o0);
    pair_hit_off[pair_base + p] = o0;
    pair_fp[2 * (pair_base + p)] = h1; pair_fp[2 * (pair_base + p) + 1] = h2;
  }
}

// ---- read-name order.  smashMEM.py walks a BAM that `samtools sort -n` has put in name order (smash_mapping.sh:23-26):
// its first-wins dedupe, the order of positions.txt and with it varbin's adjacent-duplicate rule (varbin.py:56-58) all
// follow that order, not the order the reads were mapped in.  samtools 0.1.x compares names with strnum_cmp (bam_sort.c:
// digit runs as numbers, leading zeros skipped, of two equal numbers the one with FEWER leading zeros is greater; other
// bytes by value), ties broken by mate.  The pairs' names are kept, every pair is checked against its predecessor while
// its batch is appended, and tail_finish sorts the pairs by name only if some pair was out of order.
__global__ void k_pair_name_len(BatchDev b, uint64_t n_pairs, uint32_t *__restrict__ cnt) {
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (uint64_t)gridDim.x * blockDim.x)
    cnt[p] = (uint32_t)(b.name_off[2 * p + 1] - b.name_off[2 * p]);
}
__global__ void k_pair_name_copy(BatchDev b, uint64_t n_pairs, const uint64_t *__restrict__ off, uint64_t pair_base,
                                 uint64_t *counters, uint8_t *__restrict__ blob, uint64_t *__restrict__ pair_name_off) {
  const uint64_t base = counters[CNT_NAME_BYTES];
  unsigned viol = 0;
  for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t *nm = b.names + b.name_off[2 * p];
    const uint64_t len = (uint64_t)(b.name_off[2 * p + 1] - b.name_off[2 * p]), o = base + off[p];
    for (uint64_t i = 0; i < len; ++i) blob[o + i] = nm[i];
    pair_name_off[pair_base + p] = o;
    if (p + 1 == n_pairs) pair_name_off[pair_base + n_pairs] = o + len;
    if (p) {
      const uint64_t lp = (uint64_t)(b.name_off[2 * p - 1] - b.name_off[2 * p - 2]);
      if (strnum_cmp(b.names + b.name_off[2 * p - 2], lp, nm, len) > 0) ++viol;
    } else if (pair_base) {                                   // against the last pair of the previous batch
      const uint64_t po = pair_name_off[pair_base - 1];
      if (strnum_cmp(blob + po, base - po, nm, len) > 0) ++viol;
    }
  }
  if (viol) atomicAdd((unsigned long long *)&counters[CNT_ORDER_VIOLATIONS], (unsigned long long)viol);
}
struct NameLess {
  const uint8_t *blob; const uint64_t *off;
  __host__ __device__ bool operator()(uint32_t x, uint32_t y) const {
    const int c = strnum_cmp(blob + off[x], off[x + 1] - off[x], blob + off[y], off[y + 1] - off[y]);
    return c < 0 || (c == 0 && x < y);                        // equal names keep their arrival order (stable merge sort)
  }
};
__global__ void k_perm_iota(uint32_t *__restrict__ v, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}
__global__ void k_permute_pairs(const uint32_t *__restrict__ perm, uint64_t n, const uint32_t *__restrict__ nhits, const uint64_t *__restrict__ fp,
                                const uint64_t *__restrict__ hit_off, uint32_t *__restrict__ nhits2, uint64_t *__restrict__ fp2,
                                uint64_t *__restrict__ hit_off2) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t j = perm[i];
    nhits2[i] = nhits[j]; fp2[2 * i] = fp[2 * j]; fp2[2 * i + 1] = fp[2 * j + 1]; hit_off2[i] = hit_off[j];
  }
}

// exclusive scan helper lives in kernels.cu
int exclusive_scan_u32_public(const uint32_t *in, uint64_t n, uint64_t *blk, uint64_t *out, cudaStream_t st);

__global__ void k_bump(uint64_t *total, const uint64_t *batch_total) { *total += *batch_total; }

// No host synchronisation here: the running hit total lives on the device (t->d_nhits) and the
// host only keeps an upper bound (every kept hit is one of the batch's records) for capacity.
int tail_accumulate(TailState *t, const DevIndex &ix, const BatchDev &b, const WorkDev &w,
                    uint64_t n_records_bound, uint64_t name_bytes_bound, cudaStream_t st, uint64_t *launches) {
  if (!t->configured) return tfail(SMASH_ERR_STATE, "smash_tail_configure has not been called");
  if (!ix.mapbody) return tfail(SMASH_ERR_STATE, "map.bin not loaded");
  const uint64_t n_pairs = (b.n_reads + 1) / 2;
  if (!n_pairs) return 0;
  if (!t->d_nhits) { TCU(cudaMalloc((void **)&t->d_nhits, 8 * N_TAIL_COUNTERS)); TCU(cudaMemset(t->d_nhits, 0, 8 * N_TAIL_COUNTERS)); }
  if (!t->last_ev) TCU(cudaEventCreateWithFlags(&t->last_ev, cudaEventDisableTiming));
  // batches come from SMASH_N_SLOTS streams: chain the appends so they never overlap
  if (t->ev_recorded) TCU(cudaStreamWaitEvent(st, t->last_ev, 0));
  if (n_pairs + 2 > t->batch_cap) {
    if (t->batch_cnt) cudaFree(t->batch_cnt);
    if (t->batch_off) cudaFree(t->batch_off);
    if (t->blk) cudaFree(t->blk);
    t->batch_cap = n_pairs + n_pairs / 4 + 1024;
    TCU(cudaMalloc((void **)&t->batch_cnt, 4 * t->batch_cap));
    TCU(cudaMalloc((void **)&t->batch_off, 8 * (t->batch_cap + 1)));
    TCU(cudaMalloc((void **)&t->blk, 8 * (t->batch_cap / 2048 + 8)));
  }
  int rc;
  if (t->ev_recorded && (t->n_pairs + n_pairs > t->pair_nhits.cap || t->n_hits_bound + n_records_bound + 1 > t->hits.cap ||
                         t->n_pairs + n_pairs + 1 > t->pair_name_off.cap || t->n_name_bound + name_bytes_bound + 1 > t->name_blob.cap))
    TCU(cudaEventSynchronize(t->last_ev));          // growing: the previous append must have landed
  const size_t want_pairs = t->pair_nhits.cap ? t->n_pairs + n_pairs : 8 * n_pairs;      // first batch: room for 8
  const size_t want_hits = t->hits.cap ? t->n_hits_bound + n_records_bound + 1 : 8 * (n_records_bound + 1);
  if ((rc = t->pair_nhits.reserve(want_pairs, t->n_pairs, st)) || (rc = t->pair_fp.reserve(2 * want_pairs, 2 * t->n_pairs, st)) ||
      (rc = t->pair_hit_off.reserve(want_pairs, t->n_pairs, st)) || (rc = t->hits.reserve(want_hits, t->n_hits_bound, st)))
    return rc;
  const size_t want_names = t->name_blob.cap ? t->n_name_bound + name_bytes_bound + 1 : 8 * (name_bytes_bound + 1);
  if ((rc = t->pair_name_off.reserve(want_pairs + 1, t->n_pairs + 1, st)) || (rc = t->name_blob.reserve(want_names, t->n_name_bound, st)))
    return rc;
  PairParams pp{t->hit_window, t->min_excess};
  const int grid = (int)((n_pairs + 255) / 256 < 148 * 8 ? (n_pairs + 255) / 256 : 148 * 8);
  k_pair_count<<<grid, 256, 0, st>>>(b, w, pp, n_pairs, t->batch_cnt);
  *launches += 1 + exclusive_scan_u32_public(t->batch_cnt, n_pairs, t->blk, t->batch_off, st);
  k_pair_write<<<grid, 256, 0, st>>>(b, w, pp, n_pairs, t->batch_off, t->n_pairs, t->d_nhits, t->pair_nhits.p,
                                     t->pair_fp.p, t->pair_hit_off.p, t->hits.p);
  k_bump<<<1, 1, 0, st>>>(t->d_nhits + CNT_HITS, t->batch_off + n_pairs);
  *launches += 2;
  // the pairs' read names + the name-order check (batch_cnt / batch_off are free again: same stream)
  k_pair_name_len<<<grid, 256, 0, st>>>(b, n_pairs, t->batch_cnt);
  *launches += 1 + exclusive_scan_u32_public(t->batch_cnt, n_pairs, t->blk, t->batch_off, st);
  k_pair_name_copy<<<grid, 256, 0, st>>>(b, n_pairs, t->batch_off, t->n_pairs, t->d_nhits, t->name_blob.p, t->pair_name_off.p);
  k_bump<<<1, 1, 0, st>>>(t->d_nhits + CNT_NAME_BYTES, t->batch_off + n_pairs);
  *launches += 2;
  TCU(cudaGetLastError());
  TCU(cudaEventRecord(t->last_ev, st)); t->ev_recorded = true;
  t->n_pairs += n_pairs; t->n_hits_bound += n_records_bound; t->n_name_bound += name_bytes_bound;
  return 0;
}

// ------------------------------------------------------------------ finish

constexpr uint64_t EMPTY_KEY = 0xffffffffffffffffULL;

// Table value = smallest GLOBAL pair ordinal with this key.  Local pair i has ordinal base + i;
// foreign keys (other ranks' pairs, {fp1, fp2, ordinal} triples sorted by ordinal) are inserted too.
__device__ __forceinline__ uint64_t table_key(uint64_t fp1, uint64_t fp2, uint64_t seed) {
  uint64_t h = mix64(fp1 ^ mix64(fp2 + seed));
  return h == EMPTY_KEY ? 0 : h;
}
__device__ __forceinline__ uint64_t table_claim(uint64_t *keys, uint64_t mask, uint64_t h) {
  uint64_t s = h & mask;
  for (;;) {
    const uint64_t prev = atomicCAS((unsigned long long *)&keys[s], (unsigned long long)EMPTY_KEY, (unsigned long long)h);
    if (prev == EMPTY_KEY || prev == h) return s;
    s = (s + 1) & mask;
  }
}
__global__ void k_dd_insert(const uint32_t *__restrict__ nhits, const uint64_t *__restrict__ fp, uint64_t n_pairs, uint64_t base,
                            uint64_t seed, uint64_t *keys, unsigned long long *minord, uint64_t mask, uint32_t *__restrict__ slot_of) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
    if (!nhits[i]) continue;
    const uint64_t s = table_claim(keys, mask, table_key(fp[2 * i], fp[2 * i + 1], seed));
    atomicMin(&minord[s], (unsigned long long)(base + i));
    slot_of[i] = (uint32_t)s;
  }
}
__global__ void k_dd_insert_foreign(const uint64_t *__restrict__ fk, uint64_t n, uint64_t seed, uint64_t *keys,
                                    unsigned long long *minord, uint64_t mask) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t s = table_claim(keys, mask, table_key(fk[3 * i], fk[3 * i + 1], seed));
    atomicMin(&minord[s], (unsigned long long)fk[3 * i + 2]);
  }
}
// keep[i] = 1 first occurrence of its key, 0 empty or duplicate (smashMEM.py:217-228)
__global__ void k_dd_resolve(const uint32_t *__restrict__ nhits, const uint64_t *__restrict__ fp, const uint64_t *__restrict__ hit_off,
                             const uint64_t *__restrict__ hits, uint64_t n_pairs, uint64_t base, const unsigned long long *__restrict__ minord,
                             const uint32_t *__restrict__ slot_of, const uint64_t *__restrict__ fk, uint64_t n_foreign,
                             uint8_t *__restrict__ keep, uint64_t *stats /*[0]=dupes,[1]=nondupes,[2]=unresolved*/) {
  unsigned dup = 0, non = 0, unres = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t n = nhits[i];
    if (!n) { keep[i] = 0; continue; }
    const uint64_t m = minord[slot_of[i]];
    if (m == base + i) { keep[i] = 1; ++non; continue; }
    bool same;
    if (m >= base && m < base + n_pairs) {                   // an earlier pair of this rank: compare the hit lists
      const uint64_t j = m - base;
      same = nhits[j] == n;
      for (uint32_t k = 0; same && k < n; ++k) same = hits[hit_off[j] + k] == hits[hit_off[i] + k];
    } else {                                                 // another rank's pair: compare the 128-bit fingerprint
      uint64_t lo = 0, hi = n_foreign;
      while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (fk[3 * mid + 2] < m) lo = mid + 1; else hi = mid; }
      same = lo < n_foreign && fk[3 * lo + 2] == m && fk[3 * lo] == fp[2 * i] && fk[3 * lo + 1] == fp[2 * i + 1];
    }
    keep[i] = 0;
    if (same) ++dup; else ++unres;
  }
  // one atomic per warp and counter instead of one per pair
  dup = __reduce_add_sync(0xffffffffu, dup); non = __reduce_add_sync(0xffffffffu, non); unres = __reduce_add_sync(0xffffffffu, unres);
  if ((threadIdx.x & 31) == 0) {
    if (dup) atomicAdd((unsigned long long *)&stats[0], (unsigned long long)dup);
    if (non) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)non);
    if (unres) atomicAdd((unsigned long long *)&stats[2], (unsigned long long)unres);
  }
}
// Dupe decision from a precomputed verdict: min_ord[j] = smallest global ordinal carrying the key of
// this rank's j-th exported (non-empty) pair; off[i] = export index of pair i.
__global__ void k_dd_from_verdict(const uint32_t *__restrict__ nhits, const uint64_t *__restrict__ off, const uint64_t *__restrict__ min_ord,
                                  uint64_t n_pairs, uint64_t base, uint8_t *__restrict__ keep, uint64_t *stats) {
  unsigned dup = 0, non = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
    if (!nhits[i]) { keep[i] = 0; continue; }
    const bool first = min_ord[off[i]] == base + i;
    keep[i] = first ? 1 : 0;
    if (first) ++non; else ++dup;
  }
  dup = __reduce_add_sync(0xffffffffu, dup); non = __reduce_add_sync(0xffffffffu, non);
  if ((threadIdx.x & 31) == 0) {
    if (dup) atomicAdd((unsigned long long *)&stats[0], (unsigned long long)dup);
    if (non) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)non);
  }
}
// {fp1, fp2, ordinal} of every non-empty pair, in pair order (=> sorted by ordinal)
__global__ void k_export_flags(const uint32_t *__restrict__ nhits, uint64_t n, uint32_t *__restrict__ flag) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) flag[i] = nhits[i] != 0;
}
__global__ void k_export_keys(const uint32_t *__restrict__ nhits, const uint64_t *__restrict__ fp, const uint64_t *__restrict__ off,
                              uint64_t n, uint64_t base, uint64_t *__restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (nhits[i]) { const uint64_t o = off[i]; out[3 * o] = fp[2 * i]; out[3 * o + 1] = fp[2 * i + 1]; out[3 * o + 2] = base + i; }
}
// per surviving pair: how many hits pass the regex (positions.txt) / varbin's chromosome filters
__global__ void k_pair_out_count(const uint32_t *__restrict__ nhits, const uint64_t *__restrict__ hit_off,
                                 const uint64_t *__restrict__ hits, const uint8_t *__restrict__ keep, uint64_t n_pairs,
                                 const int64_t *__restrict__ chrom_off, uint32_t *__restrict__ c_pos, uint32_t *__restrict__ c_bin) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t a = 0, bct = 0;
    if (keep[i]) {
      for (uint32_t k = 0; k < nhits[i]; ++k) {
        const int64_t o = chrom_off[hits[hit_off[i] + k] >> 40];
        a += o != -1; bct += o >= 0;
      }
    }
    c_pos[i] = a; c_bin[i] = bct;
  }
}
__global__ void k_pair_out_write(const uint32_t *__restrict__ nhits, const uint64_t *__restrict__ hit_off,
                                 const uint64_t *__restrict__ hits, const uint8_t *__restrict__ keep, uint64_t n_pairs,
                                 const int64_t *__restrict__ chrom_off, const uint64_t *__restrict__ o_pos,
                                 const uint64_t *__restrict__ o_bin, int32_t *__restrict__ pos_chrom, int64_t *__restrict__ pos_pos,
                                 int64_t *__restrict__ f_pos, int64_t *__restrict__ f_abs) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
    if (!keep[i]) continue;
    uint64_t a = o_pos[i], bq = o_bin[i];
    for (uint32_t k = 0; k < nhits[i]; ++k) {
      const uint64_t v = hits[hit_off[i] + k];
      const int64_t pos = (int64_t)(v & ((1ull << 40) - 1));
      const int64_t o = chrom_off[v >> 40];
      if (o != -1) { pos_chrom[a] = (int32_t)(v >> 40); pos_pos[a] = pos; ++a; }
      if (o >= 0) { f_pos[bq] = pos; f_abs[bq] = pos + o; ++bq; }
    }
  }
}

// bisect.bisect (right) - 1 over the bin starts (varbin.py:89-92).  With the granule table of tail_configure
// (lut[g] = number of starts <= g << shift) the search runs between lut[g] and lut[g + 1]: bins are far wider than a
// granule, so it ends after zero to two probes instead of log2(n_bins) = 16..19 dependent loads.
struct BinLut { const uint32_t *lut; uint64_t n; int shift; };
__device__ __forceinline__ uint64_t bin_of(const int64_t *__restrict__ starts, uint64_t n_bins, int64_t abspos, const BinLut &bl) {
  uint64_t lo = 0, hi = n_bins;
  if (bl.n && abspos >= 0) {
    const uint64_t g = (uint64_t)abspos >> bl.shift;
    if (g + 1 < bl.n) { lo = bl.lut[g]; hi = bl.lut[g + 1]; } else lo = bl.lut[bl.n - 1];
  }
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (starts[mid] <= abspos) lo = mid + 1; else hi = mid; }
  return lo ? lo - 1 : n_bins - 1;                           // python's counts[-1]
}
// varbin.py:56-58: a line whose position string equals the previous kept line's is a duplicate
__global__ void k_varbin_global(const int64_t *__restrict__ f_pos, const int64_t *__restrict__ f_abs, uint64_t n,
                                const int64_t *__restrict__ starts, uint64_t n_bins, unsigned long long *counts,
                                unsigned long long *stats /*[3]=dups*/, int has_prev, int64_t prev_pos, BinLut bl) {
  unsigned long long dups = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (i ? f_pos[i] == f_pos[i - 1] : (has_prev && f_pos[0] == prev_pos)) { ++dups; continue; }
    atomicAdd(&counts[bin_of(starts, n_bins, f_abs[i], bl)], 1ull);
  }
  if (dups) atomicAdd(&stats[3], dups);
}
__global__ void k_varbin_smem(const int64_t *__restrict__ f_pos, const int64_t *__restrict__ f_abs, uint64_t n,
                              const int64_t *__restrict__ starts, uint64_t n_bins, unsigned long long *counts,
                              unsigned long long *stats, int has_prev, int64_t prev_pos, BinLut bl) {
  extern __shared__ uint32_t hist[];
  for (uint64_t k = threadIdx.x; k < n_bins; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  unsigned long long dups = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (i ? f_pos[i] == f_pos[i - 1] : (has_prev && f_pos[0] == prev_pos)) { ++dups; continue; }
    atomicAdd(&hist[bin_of(starts, n_bins, f_abs[i], bl)], 1u);
  }
  if (dups) atomicAdd(&stats[3], dups);
  __syncthreads();
  for (uint64_t k = threadIdx.x; k < n_bins; k += blockDim.x) if (hist[k]) atomicAdd(&counts[k], (unsigned long long)hist[k]);
}

// Phase A: duplicate removal (local pairs + foreign keys), ordered compaction into the positions
// list and the varbin-filtered list; reports the shard's edge (first/last filtered position).
int tail_phase_a(TailState *t, uint64_t ordinal_base, const uint64_t *foreign_keys, uint64_t n_foreign,
                 smash_tail_edge *edge, cudaStream_t st, uint64_t *launches, const uint64_t *verdict_min_ord, bool sharded) {
  if (!t->configured) return tfail(SMASH_ERR_STATE, "smash_tail_configure has not been called");
  const uint64_t P = t->n_pairs;
  double tq = t_now();
  uint64_t h_stats[4] = {0, 0, 0, 0};
  uint64_t n_pos = 0, n_f = 0;
  t->a_done = false;
  if (P) {
    uint64_t tsize = 1024; while (tsize < 2 * (P + n_foreign)) tsize <<= 1;
    // persistent scratch (grown on demand, never freed per call: cudaMalloc/cudaFree cost milliseconds)
    int rcs;
    if ((rcs = t->scr[0].reserve(8 * tsize, 0, st)) || (rcs = t->scr[1].reserve(8 * tsize, 0, st)) || (rcs = t->scr[2].reserve(4 * P, 0, st)) ||
        (rcs = t->scr[3].reserve(P, 0, st)) || (rcs = t->scr[4].reserve(64, 0, st)) || (rcs = t->scr[5].reserve(4 * P, 0, st)) ||
        (rcs = t->scr[6].reserve(4 * P, 0, st)) || (rcs = t->scr[7].reserve(8 * (P + 1), 0, st)) || (rcs = t->scr[8].reserve(8 * (P + 1), 0, st)) ||
        (rcs = t->scr[9].reserve(8 * (P / 2048 + 8), 0, st)))
      return rcs;
    uint64_t *keys = (uint64_t *)t->scr[0].p, *d_stats = (uint64_t *)t->scr[4].p, *o_pos = (uint64_t *)t->scr[7].p,
             *o_bin = (uint64_t *)t->scr[8].p, *blk = (uint64_t *)t->scr[9].p;
    unsigned long long *minord = (unsigned long long *)t->scr[1].p;
    uint32_t *slot_of = (uint32_t *)t->scr[2].p, *c_pos = (uint32_t *)t->scr[5].p, *c_bin = (uint32_t *)t->scr[6].p;
    uint8_t *keep = (uint8_t *)t->scr[3].p;
    TDBG("reserve");
    const int grid = (int)((P + 255) / 256 < 148 * 8 ? (P + 255) / 256 : 148 * 8);
    // pairs in `samtools sort -n` order: the arrival order unless the append-time check saw a pair out of place
    const uint32_t *nh = t->pair_nhits.p; const uint64_t *pfp = t->pair_fp.p, *pho = t->pair_hit_off.p;
    {
      uint64_t h_cnt[N_TAIL_COUNTERS] = {0, 0, 0, 0};
      TCU(cudaMemcpyAsync(h_cnt, t->d_nhits, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
      TCU(cudaStreamSynchronize(st));
      t->order_violations = h_cnt[CNT_ORDER_VIOLATIONS];
    }
    if (t->order_violations) {
      if (sharded)
        return tfail(SMASH_ERR_DATA, "%llu read pairs are out of `samtools sort -n` name order: the read-sharded tail needs name-ordered "
                     "input (smashMEM.py sees a name-sorted BAM, smash_mapping.sh:23); use smash_tail_finish on one GPU",
                     (unsigned long long)t->order_violations);
      if (P >= 0xffffffffull) return tfail(SMASH_ERR_ARG, "too many pairs for the name sort");
      if ((rcs = t->scr[12].reserve(4 * P, 0, st)) || (rcs = t->scr[13].reserve(4 * P, 0, st)) || (rcs = t->scr[14].reserve(16 * P, 0, st)) ||
          (rcs = t->scr[15].reserve(8 * P, 0, st)))
        return rcs;
      uint32_t *perm = (uint32_t *)t->scr[12].p, *nh2 = (uint32_t *)t->scr[13].p;
      uint64_t *fp2 = (uint64_t *)t->scr[14].p, *ho2 = (uint64_t *)t->scr[15].p;
      k_perm_iota<<<grid, 256, 0, st>>>(perm, P);
      const NameLess less{t->name_blob.p, t->pair_name_off.p};
#if !defined(SMASH_CUDA_SHIM)
      size_t tmp_bytes = 0;
      TCU(cub::DeviceMergeSort::SortKeys(nullptr, tmp_bytes, perm, (int64_t)P, less, st));
      if ((rcs = t->scr[16].reserve(tmp_bytes + 16, 0, st))) return rcs;
      TCU(cub::DeviceMergeSort::SortKeys(t->scr[16].p, tmp_bytes, perm, (int64_t)P, less, st));
#else
      TCU(cudaStreamSynchronize(st));
      std::sort(perm, perm + P, less);                       // host emulation (tests/emul): device memory is host memory there
#endif
      k_permute_pairs<<<grid, 256, 0, st>>>(perm, P, t->pair_nhits.p, t->pair_fp.p, t->pair_hit_off.p, nh2, fp2, ho2);
      *launches += 3;
      nh = nh2; pfp = fp2; pho = ho2;
      TDBG("name sort");
    }
    if (verdict_min_ord) {
      // the host already resolved the global first-wins rule (multigpu.py: partitioned exchange)
      uint32_t *flag = c_pos;                       // scratch reuse: export index of every pair
      k_export_flags<<<grid, 256, 0, st>>>(nh, P, flag);
      *launches += 1 + exclusive_scan_u32_public(flag, P, blk, o_pos, st);
      TCU(cudaMemsetAsync(d_stats, 0, 32, st));
      k_dd_from_verdict<<<grid, 256, 0, st>>>(nh, o_pos, verdict_min_ord, P, ordinal_base, keep, d_stats);
      *launches += 1;
      TCU(cudaMemcpyAsync(h_stats, d_stats, 32, cudaMemcpyDeviceToHost, st));
      TCU(cudaStreamSynchronize(st));
    } else
    for (uint64_t seed = 1;; ++seed) {
      TCU(cudaMemsetAsync(keys, 0xff, 8 * tsize, st)); TCU(cudaMemsetAsync(minord, 0xff, 8 * tsize, st));
      TCU(cudaMemsetAsync(d_stats, 0, 32, st));
      const uint64_t sd = seed * 0x9e3779b97f4a7c15ULL;
      if (n_foreign) { k_dd_insert_foreign<<<grid, 256, 0, st>>>(foreign_keys, n_foreign, sd, keys, minord, tsize - 1); *launches += 1; }
      k_dd_insert<<<grid, 256, 0, st>>>(nh, pfp, P, ordinal_base, sd, keys, minord, tsize - 1, slot_of);
      k_dd_resolve<<<grid, 256, 0, st>>>(nh, pfp, pho, t->hits.p, P, ordinal_base, minord, slot_of,
                                         foreign_keys, n_foreign, keep, d_stats);
      *launches += 2;
      TCU(cudaMemcpyAsync(h_stats, d_stats, 32, cudaMemcpyDeviceToHost, st));
      TCU(cudaStreamSynchronize(st));
      if (h_stats[2] == 0) break;                            // a 64-bit table-key collision: re-key
      if (seed > 8) return tfail(SMASH_ERR_DATA, "duplicate-key table could not be resolved");
    }
    TDBG("dedupe");
    k_pair_out_count<<<grid, 256, 0, st>>>(nh, pho, t->hits.p, keep, P, t->chrom_off, c_pos, c_bin);
    *launches += 1 + exclusive_scan_u32_public(c_pos, P, blk, o_pos, st);
    *launches += exclusive_scan_u32_public(c_bin, P, blk, o_bin, st);
    TCU(cudaMemcpyAsync(&n_pos, o_pos + P, 8, cudaMemcpyDeviceToHost, st));
    TCU(cudaMemcpyAsync(&n_f, o_bin + P, 8, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
    if (n_pos + 1 > t->pos_cap) {
      if (t->pos_chrom) cudaFree(t->pos_chrom);
      if (t->pos_pos) cudaFree(t->pos_pos);
      t->pos_cap = n_pos + 1024;
      TCU(cudaMalloc((void **)&t->pos_chrom, 4 * t->pos_cap)); TCU(cudaMalloc((void **)&t->pos_pos, 8 * t->pos_cap));
    }
    TDBG("out_count+scans");
    if ((rcs = t->scr[10].reserve(8 * (n_f + 1), 0, st)) || (rcs = t->scr[11].reserve(8 * (n_f + 1), 0, st))) return rcs;
    int64_t *f_pos = (int64_t *)t->scr[10].p, *f_abs = (int64_t *)t->scr[11].p;
    k_pair_out_write<<<grid, 256, 0, st>>>(nh, pho, t->hits.p, keep, P, t->chrom_off, o_pos, o_bin,
                                           t->pos_chrom, t->pos_pos, f_pos, f_abs);
    *launches += 1;
    TDBG("out_write");
  }
  t->n_positions = n_pos; t->n_f = n_f; t->a_dupes = h_stats[0]; t->a_non_dupes = h_stats[1];
  smash_tail_edge e{}; e.n_filtered = n_f;
  if (n_f) {
    const int64_t *f_pos = (const int64_t *)t->scr[10].p;
    TCU(cudaMemcpyAsync(&e.first_pos, f_pos, 8, cudaMemcpyDeviceToHost, st));
    TCU(cudaMemcpyAsync(&e.last_pos, f_pos + (n_f - 1), 8, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
  }
  TCU(cudaGetLastError());
  if (edge) *edge = e;
  t->a_done = true;
  return 0;
}

// Phase B: varbin's adjacent-duplicate rule (optionally seeded with the previous shard's last
// filtered position) + bin histogram.
int tail_phase_b(TailState *t, int has_prev, int64_t prev_last_pos, int64_t *counts_host, int64_t *counts_device,
                 smash_tail_stats *stats, cudaStream_t st, uint64_t *launches) {
  if (!t->a_done) return tfail(SMASH_ERR_STATE, "phase A has not run");
  double tq = t_now();
  TCU(cudaMemsetAsync(t->counts, 0, 8 * t->n_bins, st));
  uint64_t h_stats[4] = {0, 0, 0, 0};
  const uint64_t n_f = t->n_f;
  if (n_f) {
    int rcs;
    if ((rcs = t->scr[4].reserve(64, 0, st))) return rcs;
    uint64_t *d_stats = (uint64_t *)t->scr[4].p;
    TCU(cudaMemsetAsync(d_stats, 0, 32, st));
    const int64_t *f_pos = (const int64_t *)t->scr[10].p, *f_abs = (const int64_t *)t->scr[11].p;
    const size_t smem = 4 * t->n_bins;
    const BinLut bl{t->bin_lut, t->lut_n, t->lut_shift};
    const int vgrid = (int)((n_f + 255) / 256 < 148 * 2 ? (n_f + 255) / 256 : 148 * 2);
    if (smem <= 200 * 1024 && n_f >= 16 * t->n_bins) {
      TCU(cudaFuncSetAttribute(k_varbin_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int g = vgrid < 148 ? vgrid : 148;
      k_varbin_smem<<<g, 1024, smem, st>>>(f_pos, f_abs, n_f, t->bin_starts, t->n_bins, (unsigned long long *)t->counts,
                                           (unsigned long long *)d_stats, has_prev, prev_last_pos, bl);
    } else {
      k_varbin_global<<<vgrid, 256, 0, st>>>(f_pos, f_abs, n_f, t->bin_starts, t->n_bins, (unsigned long long *)t->counts,
                                             (unsigned long long *)d_stats, has_prev, prev_last_pos, bl);
    }
    *launches += 1;
    TCU(cudaMemcpyAsync(h_stats, d_stats, 32, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
    TDBG("varbin");
  }
  smash_tail_stats s{};
  s.n_dupe_pairs = t->a_dupes; s.n_non_dupe_pairs = t->a_non_dupes;
  s.n_positions = t->n_positions; s.total_reads = n_f; s.dups_removed = h_stats[3]; s.reads_kept = n_f - h_stats[3];
  if (counts_device) TCU(cudaMemcpyAsync(counts_device, t->counts, 8 * t->n_bins, cudaMemcpyDeviceToDevice, st));
  if (counts_host) TCU(cudaMemcpyAsync(counts_host, t->counts, 8 * t->n_bins, cudaMemcpyDeviceToHost, st));
  TCU(cudaStreamSynchronize(st));
  TCU(cudaGetLastError());
  if (stats) *stats = s;
  return 0;
}

int tail_finish(TailState *t, int64_t *counts_host, int64_t *counts_device, smash_tail_stats *stats,
                cudaStream_t st, uint64_t *launches) {
  int rc = tail_phase_a(t, 0, nullptr, 0, nullptr, st, launches, nullptr, false);
  if (rc) return rc;
  return tail_phase_b(t, 0, 0, counts_host, counts_device, stats, st, launches);
}

// {fp1, fp2, ordinal} triples of this rank's non-empty pairs (device memory owned by the tail).
int tail_export_keys(TailState *t, uint64_t ordinal_base, const uint64_t **dev_keys, uint64_t *n, cudaStream_t st, uint64_t *launches) {
  const uint64_t P = t->n_pairs;
  *dev_keys = nullptr; *n = 0;
  if (!P) return 0;
  int rcs;
  if ((rcs = t->scr[5].reserve(4 * P, 0, st)) || (rcs = t->scr[7].reserve(8 * (P + 1), 0, st)) || (rcs = t->scr[9].reserve(8 * (P / 2048 + 8), 0, st)))
    return rcs;
  uint32_t *flag = (uint32_t *)t->scr[5].p; uint64_t *off = (uint64_t *)t->scr[7].p, *blk = (uint64_t *)t->scr[9].p;
  const int grid = (int)((P + 255) / 256 < 148 * 8 ? (P + 255) / 256 : 148 * 8);
  k_export_flags<<<grid, 256, 0, st>>>(t->pair_nhits.p, P, flag);
  *launches += 1 + exclusive_scan_u32_public(flag, P, blk, off, st);
  uint64_t cnt = 0;
  TCU(cudaMemcpyAsync(&cnt, off + P, 8, cudaMemcpyDeviceToHost, st));
  TCU(cudaStreamSynchronize(st));
  if ((rcs = t->exp_keys.reserve(24 * (cnt + 1), 0, st))) return rcs;
  k_export_keys<<<grid, 256, 0, st>>>(t->pair_nhits.p, t->pair_fp.p, off, P, ordinal_base, (uint64_t *)t->exp_keys.p);
  *launches += 1;
  TCU(cudaStreamSynchronize(st));
  TCU(cudaGetLastError());
  *dev_keys = (const uint64_t *)t->exp_keys.p; *n = cnt;
  return 0;
}

// Capacity hint: allocate everything tail_accumulate / tail_finish will need for `pairs` read pairs
// and `hits` kept hits, so that no cudaMalloc/cudaFree happens later (they cost milliseconds each).
int tail_reserve(TailState *t, uint64_t pairs, uint64_t hits, cudaStream_t st) {
  int rc;
  if ((rc = t->pair_nhits.reserve(pairs, t->n_pairs, st)) || (rc = t->pair_fp.reserve(2 * pairs, 2 * t->n_pairs, st)) ||
      (rc = t->pair_hit_off.reserve(pairs, t->n_pairs, st)) || (rc = t->hits.reserve(hits + 1, t->n_hits_bound, st)))
    return rc;
  uint64_t tsize = 1024; while (tsize < 2 * pairs) tsize <<= 1;
  if ((rc = t->pair_name_off.reserve(pairs + 2, t->n_pairs + 1, st))) return rc;
  const size_t want[12] = {8 * tsize, 8 * tsize, 4 * pairs, pairs, 64, 4 * pairs, 4 * pairs, 8 * (pairs + 1), 8 * (pairs + 1),
                           8 * (pairs / 2048 + 8), 8 * (hits + 1), 8 * (hits + 1)};
  for (int i = 0; i < 12; ++i) if ((rc = t->scr[i].reserve(want[i], 0, st))) return rc;
  if ((rc = t->exp_keys.reserve(24 * (pairs + 1), 0, st))) return rc;          // smash_tail_export_keys (multi-GPU tail)
  if (hits + 1 > t->pos_cap) {
    if (t->pos_chrom) cudaFree(t->pos_chrom);
    if (t->pos_pos) cudaFree(t->pos_pos);
    t->pos_cap = hits + 1024;
    TCU(cudaMalloc((void **)&t->pos_chrom, 4 * t->pos_cap)); TCU(cudaMalloc((void **)&t->pos_pos, 8 * t->pos_cap));
  }
  return 0;
}

int tail_positions(TailState *t, const int32_t **chrom, const int64_t **pos, uint64_t *n) {
  const uint64_t m = t->n_positions;
  if (m + 1 > t->h_pos_cap) {
    if (t->h_pos_chrom) cudaFreeHost(t->h_pos_chrom);
    if (t->h_pos_pos) cudaFreeHost(t->h_pos_pos);
    t->h_pos_cap = m + 1024;
    TCU(cudaHostAlloc((void **)&t->h_pos_chrom, 4 * t->h_pos_cap, cudaHostAllocDefault));
    TCU(cudaHostAlloc((void **)&t->h_pos_pos, 8 * t->h_pos_cap, cudaHostAllocDefault));
  }
  if (m) {
    TCU(cudaMemcpy(t->h_pos_chrom, t->pos_chrom, 4 * m, cudaMemcpyDeviceToHost));
    TCU(cudaMemcpy(t->h_pos_pos, t->pos_pos, 8 * m, cudaMemcpyDeviceToHost));
  }
  *chrom = t->h_pos_chrom; *pos = t->h_pos_pos; *n = m;
  return 0;
}

// ---- several shards' pairs gathered into one tail (comm.cu).  The read-sharded protocol needs every shard to be a
// contiguous run of the name order; when the input is not name-sorted the shards' pairs are appended to rank 0's tail in
// rank order (= submission order of the whole input) and the single-GPU finish, with its name sort, does the rest.
int tail_totals(TailState *t, uint64_t out[4], cudaStream_t st) {
  uint64_t h[N_TAIL_COUNTERS] = {0, 0, 0, 0};
  if (t->d_nhits) {
    TCU(cudaMemcpyAsync(h, t->d_nhits, sizeof h, cudaMemcpyDeviceToHost, st));
    TCU(cudaStreamSynchronize(st));
  }
  out[0] = t->n_pairs; out[1] = h[CNT_HITS]; out[2] = h[CNT_NAME_BYTES]; out[3] = h[CNT_ORDER_VIOLATIONS];
  return 0;
}
int tail_edge_names(TailState *t, uint8_t *first, uint8_t *last, cudaStream_t st) {
  memset(first, 0, TAIL_EDGE_NAME); memset(last, 0, TAIL_EDGE_NAME);
  const uint64_t P = t->n_pairs;
  if (!P) return 0;
  uint64_t o[4];
  TCU(cudaMemcpyAsync(o, t->pair_name_off.p, 16, cudaMemcpyDeviceToHost, st));
  TCU(cudaMemcpyAsync(o + 2, t->pair_name_off.p + (P - 1), 16, cudaMemcpyDeviceToHost, st));
  TCU(cudaStreamSynchronize(st));
  uint8_t *dst[2] = {first, last};
  for (int k = 0; k < 2; ++k) {
    const uint64_t len = o[2 * k + 1] - o[2 * k];
    if (len >= 255) { dst[k][0] = 255; continue; }
    dst[k][0] = (uint8_t)len;
    if (len) TCU(cudaMemcpyAsync(dst[k] + 1, t->name_blob.p + o[2 * k], len, cudaMemcpyDeviceToHost, st));
  }
  TCU(cudaStreamSynchronize(st));
  return 0;
}
// room for `pairs`/`hits`/`name_bytes` MORE than the tail holds now (exact totals: the appends have all landed)
int tail_absorb_reserve(TailState *t, uint64_t pairs, uint64_t hits, uint64_t name_bytes, cudaStream_t st) {
  uint64_t cur[4];
  int rc;
  if ((rc = tail_totals(t, cur, st))) return rc;
  if (!t->d_nhits) { TCU(cudaMalloc((void **)&t->d_nhits, 8 * N_TAIL_COUNTERS)); TCU(cudaMemset(t->d_nhits, 0, 8 * N_TAIL_COUNTERS)); }
  const uint64_t P = cur[0] + pairs;
  if ((rc = t->pair_nhits.reserve(P, cur[0], st)) || (rc = t->pair_fp.reserve(2 * P, 2 * cur[0], st)) ||
      (rc = t->pair_hit_off.reserve(P, cur[0], st)) || (rc = t->hits.reserve(cur[1] + hits + 1, cur[1], st)) ||
      (rc = t->pair_name_off.reserve(P + 2, cur[0] ? cur[0] + 1 : 0, st)) || (rc = t->name_blob.reserve(cur[2] + name_bytes + 1, cur[2], st)))
    return rc;
  if (!cur[0]) TCU(cudaMemsetAsync(t->pair_name_off.p, 0, 8, st));
  return 0;
}
__global__ void k_absorb_fixup(uint64_t *__restrict__ hit_off, uint64_t *__restrict__ name_off, uint64_t from, uint64_t n, uint64_t hit_base,
                               uint64_t name_base) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    hit_off[from + i] += hit_base;
    name_off[from + i + 1] += name_base;                      // the shard's offsets 1..n arrived; name_off[from] is this tail's old end
  }
}
__global__ void k_set_counters(uint64_t *c, uint64_t hits, uint64_t name_bytes, uint64_t violations) {
  c[CNT_HITS] = hits; c[CNT_NAME_BYTES] = name_bytes; c[CNT_ORDER_VIOLATIONS] = violations;
}
// a shard's arrays have been written behind this tail's own (pair_hit_off and pair_name_off[1..pairs] still relative to
// the shard):
// shift them, count the shard in.  The gathered pairs are by construction not in name order: the finish sorts them.
int tail_absorb_commit(TailState *t, uint64_t pairs, uint64_t hits, uint64_t name_bytes, cudaStream_t st, uint64_t *launches) {
  uint64_t cur[4];
  int rc;
  if ((rc = tail_totals(t, cur, st))) return rc;
  if (pairs) {
    const int grid = (int)((pairs + 256) / 256 < 148 * 8 ? (pairs + 256) / 256 : 148 * 8);
    k_absorb_fixup<<<grid, 256, 0, st>>>(t->pair_hit_off.p, t->pair_name_off.p, cur[0], pairs, cur[1], cur[2]);
    ++*launches;
  }
  k_set_counters<<<1, 1, 0, st>>>(t->d_nhits, cur[1] + hits, cur[2] + name_bytes, cur[3] + 1);
  ++*launches;
  TCU(cudaStreamSynchronize(st));
  TCU(cudaGetLastError());
  t->n_pairs = cur[0] + pairs; t->n_hits_bound = cur[1] + hits; t->n_name_bound = cur[2] + name_bytes;
  return 0;
}

}  // namespace smash
