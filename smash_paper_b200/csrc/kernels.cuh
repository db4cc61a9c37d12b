// kernels.cuh -- device-side views shared by kernels.cu (kernels + launchers) and api.cu (host).
#pragma once
#include <cuda_runtime.h>
#include "core.cuh"
#include "records.cuh"
#include "compact.h"

namespace smash {

struct BatchDev {
  uint64_t n_reads;
  const uint8_t *names; const int64_t *name_off;
  const uint8_t *seq; const uint8_t *qual; const int64_t *seq_off;
  const uint8_t *opt; const int64_t *opt_off;      // opt may be null
  const uint16_t *read_flag;
};

constexpr int SURV_CAP = 32;
enum { FLAG_OVERFLOW = 0, FLAG_MAPERR = 1, FLAG_LONGREAD = 2, FLAG_MAXCNT = 3, FLAG_LONGQ = 4, FLAG_SORTDUP = 5, N_FLAGS = 8 };

struct WorkDev {
  int cap;                     // slots per read (matches / items / records) when slot_off is null
  const uint64_t *slot_off;    // MEM mode: per-read slot ranges (CSR, n_reads + 1); null => read * cap
  uint64_t slots_total;        // number of slots in the batch (n_reads * cap, or slot_off[n_reads])
  Aln *aln_scratch;            // MEM mode: per-slot scratch for k_rec_build_big
  uint16_t *ord_scratch;
  Match *match_slots;          // n_reads * cap
  uint32_t *match_cnt;         // n_reads (true count, may exceed cap => overflow flag)
  Item *item_slots;            // n_reads * cap
  Rec *rec_slots;              // n_reads * cap
  ReadSum *sums;               // n_reads
  uint32_t *nrec;              // n_reads: records printed for the read
  uint64_t *rec_base;          // n_reads + 1: exclusive scan of nrec (flat record index)
  uint32_t *rec_read;          // per flat record: its read
  uint32_t *rec_bytes;         // per flat record: bytes of its SAM line
  uint64_t *rec_off;           // per flat record (+1): exclusive scan of rec_bytes = offset in `sam`
  uint64_t *sam_total;         // [0] = total SAM bytes of the batch
  uint64_t *blk_sums;          // scan scratch (n_reads / 2048 + 8)
  uint64_t *blk_sums2;         // scan scratch (n_reads * cap / 2048 + 8)
  char *sam;                   // output text
  uint64_t sam_cap;
  uint32_t *flags;             // N_FLAGS counters
  // split search (K1a/K1b): candidates that pass the seed + 4+4 filters are parked per read and verified
  // by k_mam_verify with lanes = candidates of ONE read (null => verification inside k_mam_search)
  uint64_t *surv;              // n_reads * SURV_CAP: anchor offset << 48 | left extension (0..7 exact, 8 = at least 8) << 40 | SA index
  uint8_t *surv_cnt;           // n_reads
  uint8_t *slow;               // n_reads: 1 = k_mam_seed left the read to k_mam_search (null => k_mam_search takes every read)
  uint8_t *lc;                 // lower-cased reads with the 16-byte pads of the staging buffer: read r at seq_off[r] + 32 r + 16
  // K5 record_sort (OutputSorter::flush, query.cpp:448-468): sort keys and permutation of the batch's flat records
  uint64_t *sort_abs; uint8_t *sort_flag; uint32_t *sort_perm; uint32_t *sort_bytes; uint64_t *sort_off; void *sort_tmp; size_t sort_tmp_bytes;
  // compact transport (compact.h): null => the full SAM text is written on the device (k_emit_text + k_emit_copy)
  uint32_t *cmp_bytes;         // per flat record: head + tags + lr-tags + newline
  uint64_t *cmp_off;           // per flat record (+1): exclusive scan of cmp_bytes = offset in `cmp`; total -> sam_total[1]
  char *cmp;                   // compact text of the range
  CmpMeta *cmeta;              // per flat record
  uint64_t sam_base;           // offset of the range's first line in the batch's SAM text (added to rec_off in CmpMeta)
  uint8_t *long_scratch;       // per-warp staging for reads longer than MAXQ_FAST (null if the batch has none)
  int long_q;                  // longest read of the batch
};

HD uint64_t slot_base(const WorkDev &w, uint64_t read) { return w.slot_off ? w.slot_off[read] : read * (uint64_t)w.cap; }
HD int slot_cap(const WorkDev &w, uint64_t read) {
  if (!w.slot_off) return w.cap;
  const uint64_t c = w.slot_off[read + 1] - w.slot_off[read];
  return c > 65535 ? 65535 : (int)c;
}

struct SearchParams {
  uint32_t L;                  // effective min_len = max(min_len, 2)
  int k;                       // seed length used = min(seed_k, L)
  int s;                       // anchor stride = L - k + 1
  int nucleotides_only;
  int nomap;
  int tag_mappability;
  int fast_ok;                 // 0 => every read takes the exact per-start path
  int mum;                     // -mum: cleanMUMcand sweep over the MAM matches (longSA.cpp:549-585)
};

// launchers (all asynchronous on `st`); each returns the number of kernels it launched
int launch_uniq_build(const DevIndex &ix, uint8_t *uniq, cudaStream_t st);
int launch_seed_build(const DevIndex &ix, void *seed, int k, int seed_w, cudaStream_t st);
int launch_ext_build(const DevIndex &ix, int k, uint32_t *ext, cudaStream_t st);
// 8-byte seed table -> blocked form, in place (core.cuh): count the blocks that need a row of exact values, then convert
int launch_seed_irr_count(const void *seed, uint64_t n_blocks, unsigned long long *n_irr, cudaStream_t st);
int launch_seed_block(void *seed, uint64_t n_blocks, const uint32_t *ext, uint64_t N, uint64_t *irr, unsigned long long *n_irr, cudaStream_t st);
int launch_alpha(const uint8_t *text, uint64_t N, uint32_t *alpha8, cudaStream_t st);
int launch_mam_search(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st);
int launch_mam_verify(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st);
// exact per-start MAM search with CSR slots (any number of matches per read): cnt != null => count pass into cnt,
// cnt == null => write pass (w.slot_off set; w.long_scratch sized for the longest read of the batch, w.long_q)
int launch_mam_exact(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, uint32_t *cnt, cudaStream_t st);
constexpr int MEM_STAGE = 8;     // matches per read the MEM counting pass keeps (k_mem_search): such reads are searched once
int launch_mem_count(const DevIndex &ix, const BatchDev &b, const SearchParams &p, uint32_t min_len_raw, uint32_t *cnt, Match *stage,
                     cudaStream_t st);
int launch_mem_write(const DevIndex &ix, const BatchDev &b, const SearchParams &p, uint32_t min_len_raw, const uint64_t *off,
                     Match *matches, const Match *stage, const uint32_t *cnt, cudaStream_t st);
// slot_off = exclusive scan of (match_cnt + 1): one spare slot per read for the unmapped placeholder
int launch_slot_offsets(const uint32_t *match_cnt, uint64_t n_reads, uint32_t *tmp, uint64_t *blk, uint64_t *slot_off, cudaStream_t st);
int launch_records(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st);
// host_small: [0] SAM bytes, [1..4] flags, [8] records, [9] compact bytes (sam_total[1])
int launch_publish(uint64_t *host_small, const uint64_t *sam_total, const uint64_t *rec_total, const uint32_t *flags, cudaStream_t st);
int launch_sizes_scan(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st);
int launch_emit_text(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st, uint64_t n_records);
// compact transport: head | tags | lr-tags \n of every record + its CmpMeta (no name / SEQ / QUAL bytes are written)
int launch_emit_compact(const DevIndex &ix, const BatchDev &b, const WorkDev &w, const SearchParams &p, cudaStream_t st, uint64_t n_records);
int launch_emit_copy(const BatchDev &b, const WorkDev &w, cudaStream_t st, uint64_t n_records);
// K5: rec_off re-derived so that the emit kernels write the records in MemSam::operator< order (memsam.h:136-158);
// n_records is the host-known record count.  tmp_bytes_needed != null: only report the sort's scratch size.
int launch_record_sort(const DevIndex &ix, const BatchDev &b, const WorkDev &w, uint64_t n_records, cudaStream_t st, size_t *tmp_bytes_needed);
// matches -> CSR (offsets int64[n+1] + smash_match-compatible {u64 ref, u64 query, u64 len})
int launch_match_csr(const BatchDev &b, const WorkDev &w, int64_t *off, uint64_t *triples, uint64_t *scratch, cudaStream_t st);
int launch_mappability(const DevIndex &ix, uint64_t *min_len_scratch, uint8_t *body, cudaStream_t st);

}  // namespace smash
