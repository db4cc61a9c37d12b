// tail.cuh -- device state of the smashMEM.py filter + varbin.py counting (K7/K8).
#pragma once
#include <cuda_runtime.h>
#include "../../include/smash_b200.h"
#include "kernels.cuh"

namespace smash {

template <class T> struct DGrow {       // growable device array that keeps its contents
  T *p = nullptr; size_t cap = 0;
  int reserve(size_t n, size_t used, cudaStream_t st);
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct TailState {
  bool configured = false;
  // configuration
  int64_t *bin_starts = nullptr; uint64_t n_bins = 0;
  uint32_t *bin_lut = nullptr; uint64_t lut_n = 0; int lut_shift = 0;   // granule table of bin_of (tail.cu)
  int64_t *chrom_off = nullptr; uint64_t n_chrom = 0;   // >=0 abs offset, -1 filtered by the regex, -2 by varbin
  int64_t hit_window = 10000; int32_t min_excess = 4;
  // accumulated over batches (pair index = order of submission)
  uint64_t n_pairs = 0, n_hits_bound = 0;   // hits: host-side upper bound, exact total in *d_nhits
  uint64_t *d_nhits = nullptr;
  cudaEvent_t last_ev = nullptr; bool ev_recorded = false;
  DGrow<uint32_t> pair_nhits;       // kept hits of the pair (0 => pair never reaches the dupe set)
  DGrow<uint64_t> pair_fp;          // 2 per pair: 128-bit fingerprint of the dupe key
  DGrow<uint64_t> pair_hit_off;     // first hit of the pair in `hits`
  DGrow<uint64_t> hits;             // tid << 40 | 0-based pos, r1 hits in HI order then r2 hits
  // read names of the pairs (name of read 2p), for the `samtools sort -n` order smashMEM.py sees (smash_mapping.sh:23)
  DGrow<uint8_t> name_blob; DGrow<uint64_t> pair_name_off;   // n_pairs + 1 offsets into name_blob
  uint64_t n_name_bound = 0;        // host-side upper bound of the bytes in name_blob (exact total in d_nhits[1])
  uint64_t order_violations = 0;    // pairs whose name sorts before their predecessor's (as of the last phase A)
  // scratch
  DGrow<uint8_t> scr[17];           // tail_finish work buffers (persistent)
  DGrow<uint8_t> exp_keys;          // exported {fp1,fp2,ordinal} triples
  bool a_done = false; uint64_t n_f = 0, a_dupes = 0, a_non_dupes = 0;
  uint32_t *batch_cnt = nullptr; uint64_t *batch_off = nullptr; uint64_t *blk = nullptr; size_t batch_cap = 0;
  // results of finish
  int64_t *counts = nullptr;        // n_bins
  int32_t *pos_chrom = nullptr; int64_t *pos_pos = nullptr; uint64_t n_positions = 0; size_t pos_cap = 0;
  int32_t *h_pos_chrom = nullptr; int64_t *h_pos_pos = nullptr; size_t h_pos_cap = 0;
};

// samtools 0.1.x name order (bam_sort.c strnum_cmp), shared by the append-time order check, the name sort and comm.cu
__host__ __device__ inline bool nm_digit(const uint8_t *s, uint64_t n, uint64_t i) { return i < n && s[i] >= '0' && s[i] <= '9'; }
__host__ __device__ inline int strnum_cmp(const uint8_t *a, uint64_t na, const uint8_t *b, uint64_t nb) {
  uint64_t pa = 0, pb = 0;
  while (pa < na && pb < nb) {
    if (nm_digit(a, na, pa) && nm_digit(b, nb, pb)) {
      while (pa < na && a[pa] == '0') ++pa;
      while (pb < nb && b[pb] == '0') ++pb;
      while (nm_digit(a, na, pa) && nm_digit(b, nb, pb) && a[pa] == b[pb]) { ++pa; ++pb; }
      if (nm_digit(a, na, pa) && nm_digit(b, nb, pb)) {
        uint64_t i = 0;
        while (nm_digit(a, na, pa + i) && nm_digit(b, nb, pb + i)) ++i;
        return nm_digit(a, na, pa + i) ? 1 : nm_digit(b, nb, pb + i) ? -1 : (int)a[pa] - (int)b[pb];
      }
      if (nm_digit(a, na, pa)) return 1;
      if (nm_digit(b, nb, pb)) return -1;
      if (pa != pb) return pa < pb ? 1 : -1;
    } else {
      if (a[pa] != b[pb]) return (int)a[pa] - (int)b[pb];
      ++pa; ++pb;
    }
  }
  return pa < na ? 1 : pb < nb ? -1 : 0;
}

void tail_init(TailState *t);
void tail_release(TailState *t);
void tail_reset(TailState *t);
const char *tail_error();
int tail_configure(TailState *t, const int64_t *bin_starts, uint64_t n_bins, const int64_t *chrom_off,
                   uint64_t n_chrom, int64_t hit_window, int32_t min_excess);
int tail_accumulate(TailState *t, const DevIndex &ix, const BatchDev &b, const WorkDev &w,
                    uint64_t n_records_bound, uint64_t name_bytes_bound, cudaStream_t st, uint64_t *launches);
int tail_finish(TailState *t, int64_t *counts_host, int64_t *counts_device, smash_tail_stats *stats,
                cudaStream_t st, uint64_t *launches);
int tail_phase_a(TailState *t, uint64_t ordinal_base, const uint64_t *foreign_keys, uint64_t n_foreign,
                 smash_tail_edge *edge, cudaStream_t st, uint64_t *launches, const uint64_t *verdict_min_ord, bool sharded);
int tail_phase_b(TailState *t, int has_prev, int64_t prev_last_pos, int64_t *counts_host, int64_t *counts_device,
                 smash_tail_stats *stats, cudaStream_t st, uint64_t *launches);
int tail_export_keys(TailState *t, uint64_t ordinal_base, const uint64_t **dev_keys, uint64_t *n, cudaStream_t st, uint64_t *launches);
int tail_reserve(TailState *t, uint64_t pairs, uint64_t hits, cudaStream_t st);
int tail_positions(TailState *t, const int32_t **chrom, const int64_t **pos, uint64_t *n);
// ---- gather of several shards' pairs into one tail (comm.cu: inputs that are not in name order) ----
constexpr int TAIL_EDGE_NAME = 256;                               // bytes per edge name slot: [0] = length (255 = too long), then the name
int tail_totals(TailState *t, uint64_t out[4] /*pairs, kept hits, name bytes, pairs out of order*/, cudaStream_t st);
int tail_edge_names(TailState *t, uint8_t *first, uint8_t *last, cudaStream_t st);   // host buffers of TAIL_EDGE_NAME bytes
int tail_absorb_reserve(TailState *t, uint64_t pairs, uint64_t hits, uint64_t name_bytes, cudaStream_t st);
int tail_absorb_commit(TailState *t, uint64_t pairs, uint64_t hits, uint64_t name_bytes, cudaStream_t st, uint64_t *launches);

}  // namespace smash
