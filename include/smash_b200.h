/* include/smash_b200.h -- C ABI of libsmash_b200.so (hand-written sm_100a CUDA behind it).
 *
 * The reference (yamrom/smash-paper) has no plugin/FFI interface; its boundary is the `mummer`
 * process + the `<ref>.bin/` index files + `mapout/<chunk>.txt`, and inside the process the seam
 * `Aligner::run -> longSA::MAM|MEM(Aligner&) -> Aligner::process_match` (query.cpp:322-329,
 * 436-438).  Every entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only; all calls return 0 on success or a negative smash_status, with the
 * message available from smash_last_error().  There is NO CPU fallback: every compute call
 * fails with SMASH_ERR_CUDA when no sm_100 device / driver is present.
 *
 * Threading: one smash_ctx per GPU, calls on one ctx are serialised by the caller; different
 * ctxs (different GPUs / processes) are independent.  The index object is immutable.
 */
#ifndef SMASH_B200_H_
#define SMASH_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SMASH_OK = 0,
  SMASH_ERR_ARG = -1,      /* bad argument */
  SMASH_ERR_IO = -2,       /* index file missing / malformed (paa::Error in fasta.cpp:103-119, longSA.cpp:115-125) */
  SMASH_ERR_CUDA = -3,     /* CUDA failure or no device */
  SMASH_ERR_NOMEM = -4,
  SMASH_ERR_STATE = -5,    /* call order (e.g. bins not loaded) */
  SMASH_ERR_DATA = -6      /* input the reference itself rejects (e.g. mappability_tag.cpp:107-113) */
} smash_status;

typedef struct smash_index smash_index;   /* host view of <fa>.bin/ (mmap) or of caller arrays */
typedef struct smash_ctx smash_ctx;       /* one GPU: index in HBM + batch buffers + tail state */

/* match_t (longSA.h:78-92) */
typedef struct { uint64_t ref, query, len; } smash_match;

enum { SMASH_MODE_MUM = 0, SMASH_MODE_MAM = 1, SMASH_MODE_MEM = 2 };   /* mum_t, query.h:126 */

const char *smash_last_error(void);
/* Number of usable sm_100 devices (0 without a driver/GPU). */
int smash_device_count(void);

/* ---- index: replaces Sequence::Sequence load branch (fasta.cpp:106-137) and the longSA load
 * branch (longSA.cpp:112-136, util.cpp:100-125).  File names and formats unchanged:
 * <fa>.bin/rc{r}.ref.bin, .ref.seq.bin, rc{r}.i{4|8}.index.bin, .sa.bin, .isa.bin, .lcp.vec.bin,
 * .lcp.m.bin (+ optional map.bin).  The FASTA itself is only stat()ed for the size guard. */
int smash_index_open(const char *ref_fasta, int rcref, smash_index **out);
/* Same object from caller-owned arrays (no files): text N bytes; sa/isa N*w bytes (isa may be
 * NULL: only MEM mode and mappability generation need it); lcp_vec N bytes; lcp_m n_m 16-byte
 * items {u64 idx; u64 val}.  Pointers must stay valid until smash_index_close. */
int smash_index_from_arrays(const uint8_t *text, uint64_t N, const void *sa, const void *isa, int w,
                            const uint8_t *lcp_vec, const void *lcp_m, uint64_t n_m,
                            uint64_t n_descr, const uint64_t *startpos, const uint64_t *sizes,
                            const char *const *descr, int rcref, smash_index **out);
void smash_index_close(smash_index *ix);
uint64_t smash_index_text_len(const smash_index *ix);
int smash_index_int_width(const smash_index *ix);
/* Sequence::sam_header() (fasta.cpp:243-252); returns bytes needed. */
size_t smash_index_sam_header(const smash_index *ix, char *buf, size_t cap);

/* ---- context: replaces Pairs/Pair/Aligner construction (query.cpp:471-475, 537-563) */
typedef struct {
  int device;            /* CUDA ordinal */
  int mode;              /* SMASH_MODE_MAM (default, -mumreference) or SMASH_MODE_MEM (-maxmatch) */
  uint32_t min_len;      /* -l, default 20 (query.h:129) */
  int nomap;             /* -nomap */
  int nucleotides_only;  /* -n (query.cpp:131-138) */
  int tag_mappability;   /* append mappability_tag's L<i>/R<i> tags to every record (needs map.bin) */
  uint64_t max_batch_reads;   /* capacity hint, 0 = default */
  int seed_k;            /* 0 = auto; k-mer length of the derived seed table (<=16) */
} smash_params;
void smash_params_default(smash_params *p);

int smash_ctx_create(const smash_index *ix, const smash_params *p, smash_ctx **out);
/* Same context, but the index is BUILT on the GPU from the text (replaces the longSA build branch,
 * longSA.cpp:137-176: qsufsort + Kasai) and stays in HBM.  text = Sequence layout (fasta.cpp:151-203).
 * keep_isa: keep the inverse suffix array resident (needed for MEM mode, mappability, saving).
 * chunk_cap: suffixes sorted per pass (0 = default), bounds the builder's scratch memory. */
int smash_ctx_create_from_text(const uint8_t *text, uint64_t N, uint64_t n_descr, const uint64_t *startpos,
                               const uint64_t *sizes, const char *const *descr, int rcref, int w,
                               int keep_isa, uint64_t chunk_cap, const smash_params *p, smash_ctx **out);
/* Copy the index arrays out of HBM (any pointer may be NULL); lcp_m gets n_m .lcp.m.bin items. */
int smash_ctx_copy_index(smash_ctx *ctx, void *sa, void *isa, uint8_t *lcp_vec, void *lcp_m, uint64_t *n_m);
/* Write <ref_fasta>.bin/rc{r}.* (and map.bin) in the reference's formats (fasta.cpp:215-236,
 * longSA.cpp:179-190): what `mummer -rcref <fa> dummy` leaves behind (index_setup.sh:19,22). */
int smash_ctx_save_index(smash_ctx *ctx, const char *ref_fasta, int with_mappability);
/* Free the inverse suffix array in HBM (only MEM mode, mappability build and index save need it). */
int smash_ctx_drop_isa(smash_ctx *ctx);
void smash_ctx_destroy(smash_ctx *ctx);
/* map.bin (longSA::show_mappability output, longSA.cpp:612-690; 2 junk bytes + 2 bytes/base) for
 * the L/R tags and the smashMEM excess-mappability filter.  body = file contents after the two
 * junk bytes.  Without it smash_map_batch still works (untagged SAM), the tail does not. */
int smash_ctx_load_mappability(smash_ctx *ctx, const uint8_t *body, uint64_t n_bytes);
/* Build map.bin's body on the GPU from SA/ISA/LCP (needs isa). Copies to host if body != NULL. */
int smash_ctx_build_mappability(smash_ctx *ctx, uint8_t *body, uint64_t cap_bytes);

/* Switch the L<i>/R<i> tagging of smash_params.tag_mappability on or off between batches (mappability_tag is a separate
 * stage of smash_mapping.sh:23: `mummer`'s own mapout has no such tags).  Needs map.bin when switched on. */
int smash_ctx_set_tag_mappability(smash_ctx *ctx, int on);

/* ---- one batch of reads = what QueryReader::run hands to Pair::run (query.cpp:614-687, 481-520).
 * Reads 2k and 2k+1 are mates-by-arrival; n_reads must be even except for the last batch.
 * names: SAM column 1 without the :0/:1 suffix; read_flag: 0/65/129 as Aligner::reset derives it
 * (query.cpp:185-201); seq: original-case bases; qual; opt: "\tTAG:..." for every extra field. */
typedef struct {
  uint64_t n_reads;
  const uint8_t *names;  const int64_t *name_off;   /* n_reads+1 */
  const uint8_t *seq;    const uint8_t *qual;  const int64_t *seq_off;
  const uint8_t *opt;    const int64_t *opt_off;    /* may be NULL */
  const uint16_t *read_flag;
  uint64_t first_pair_ordinal;   /* informational only (kept for layout compatibility): the tail orders read pairs the way
                                  * the reference pipeline does -- by read name, `samtools sort -n` order (smash_mapping.sh:23) --
                                  * which is the order of submission whenever the input is already name-ordered */
} smash_batch;

typedef struct {
  uint64_t n_reads, n_matches, n_records, sam_bytes;
  /* Pointers into ctx-owned PINNED host memory, valid until the next smash_map_batch on this
   * ctx.  sam = record lines in input order (read 0's records in HI order, read 1's, ...). */
  const char *sam;
  const int64_t *match_off;       /* n_reads+1; NULL unless want_matches */
  const smash_match *matches;
  float gpu_ms;                   /* device time of this batch's kernels (CUDA events) */
} smash_result;

/* SMASH_WANT_SORTED (with SMASH_WANT_SAM): the batch is one chunk of the reference's OutputSorter -- its lines come back
 * in the order OutputSorter::flush writes them (query.cpp:448-468): sorted by MemSam::operator< (memsam.h:136-158), i.e.
 * (MemSam::chromosomes[RNAME] + POS, name, flag & (64|128|16)); lines equal in all three fail with SMASH_ERR_DATA where
 * the reference throws "flags equal".  Without it the lines are in input order (read 0's records in HI order, ...). */
enum { SMASH_WANT_SAM = 1, SMASH_WANT_MATCHES = 2, SMASH_WANT_TAIL = 4, SMASH_WANT_SORTED = 8 };

/* Pinned host memory for batches (cudaHostAlloc): buffers handed to smash_map_batch /
 * smash_submit from here are copied with true asynchronous DMA. */
void *smash_host_alloc(size_t bytes);
void smash_host_free(void *p);

/* Host buffers in, host buffers out: H2D copy, search, records, SAM text, (tail accumulation),
 * D2H copy.  Replaces Aligner::run + set_mate + print_matches for the whole batch. */
int smash_map_batch(smash_ctx *ctx, const smash_batch *b, int want, smash_result *res);

/* Double-buffered form of smash_map_batch: a ctx owns SMASH_N_SLOTS independent slots (device
 * buffers + pinned result buffers + streams + one host worker thread).  submit() hands the batch to
 * the slot's worker and returns at once; the worker enqueues H2D + kernels + D2H (this is where the
 * reference's Pair::run workers sit, query.cpp:481-520); wait() blocks until that slot's result is in
 * host memory and reports any error of the batch.  The batch's host buffers must stay valid until
 * wait().  Batches are appended to the tail in the order they were submitted, whichever slot they use.
 *
 * Transport of the SAM text: a line is NAME head SEQ QUAL tags [optional fields] L/R-tags (print_matches,
 * query.cpp:331-415), and NAME/SEQ/QUAL/optional fields are the caller's own bytes.  By default only the
 * text the GPU computes (head, tags, L/R tags: ~40 % of the line) plus 32 bytes per record cross PCIe and
 * the library's host threads rebuild the byte-identical lines from the submitted batch ("compact
 * transport", the `-qthreads` workers of the reference become these threads).  smash_submit_text and the
 * device-resident calls always produce the whole text on the device. */
#define SMASH_N_SLOTS 4
int smash_submit(smash_ctx *ctx, int slot, const smash_batch *b, int want);
int smash_wait(smash_ctx *ctx, int slot, smash_result *res);

/* Device-resident variant used to time the kernels alone: upload once, run many times. */
int smash_batch_upload(smash_ctx *ctx, const smash_batch *b);
int smash_map_resident(smash_ctx *ctx, int want, smash_result *res);  /* no H2D/D2H; res->sam NULL */
/* Copy the resident run's SAM text to host (for checks). */
int smash_fetch_sam(smash_ctx *ctx, const char **sam, uint64_t *n_bytes);

/* ---- input side on the device: raw text in, the reader's parse runs on the GPU.
 * SMASH_TEXT_SAM replaces QueryReader::run's SAM branch (query.cpp:625-648: getline, whitespace-separated
 * fields name flag <7 ignored> seq errors [optional...], ":0"/":1" from flag 64/128) together with
 * NewQuery::add_optional (query.cpp:150-153) and Aligner::reset (query.cpp:185-201).
 * SMASH_TEXT_FASTQ_PAIR takes the two (decompressed) FASTQ texts of a mate pair and replaces
 * `fastqs_to_sam fq1 fq2 [1]` (fastqs_to_sam.cpp:47-95: records alternate between the files, blank lines
 * before '@' and '+' are skipped, second header token -> XO:Z:, '>' records reuse the bases as errors,
 * empty reads print nothing, flags 77/141) followed by that SAM branch -- the producer/consumer pair of
 * smash_mapping.sh:19.  SMASH_TEXT_REPLACE_N = fastqs_to_sam's third argument (N -> Z in the bases).
 * Without SMASH_TEXT_FINAL the text is one chunk of a longer stream: only complete lines (FASTQ: complete
 * record pairs, less the last one, which may be cut) are taken, and an odd trailing read is left for the
 * next chunk so that the reader's pairing by arrival parity (query.cpp:629-637) is unchanged; `consumed`
 * says how many bytes of each text were used -- the caller passes the rest again, followed by more input.  A chunk of
 * a FASTQ pair may end after a mate-1 record: `mate2_first_next` is then 1 and the caller sets
 * SMASH_TEXT_MATE2_FIRST on the next call, so that the alternation of the two files goes on where it stopped.
 * Input the reference would misparse silently (fewer than 11 SAM fields, a non-numeric flag, SEQ/QUAL of
 * different lengths, a truncated FASTQ record) fails with SMASH_ERR_DATA, the FASTQ '@' / '+' checks with
 * the reference's own messages.  The text buffers may be reused as soon as the call returns. */
enum { SMASH_TEXT_SAM = 0, SMASH_TEXT_FASTQ_PAIR = 1 };
enum { SMASH_TEXT_FINAL = 1, SMASH_TEXT_REPLACE_N = 2, SMASH_TEXT_MATE2_FIRST = 4 };
typedef struct {
  int kind;                      /* SMASH_TEXT_* */
  int flags;                     /* SMASH_TEXT_FINAL | SMASH_TEXT_REPLACE_N | SMASH_TEXT_MATE2_FIRST */
  const char *text[2];           /* SAM: text[0]; FASTQ pair: mate-1 text, mate-2 text (host memory, pinned or not) */
  uint64_t n_bytes[2];
  uint64_t first_pair_ordinal;   /* as in smash_batch */
} smash_text;
typedef struct { uint64_t n_reads; uint64_t consumed[2]; int mate2_first_next; } smash_text_info;
/* smash_submit with the batch parsed on the device; collect the result with smash_wait. */
int smash_submit_text(smash_ctx *ctx, int slot, const smash_text *t, int want, smash_text_info *info);
/* Device-resident variant (slot 0): parse only; run it with smash_map_resident. */
int smash_text_upload(smash_ctx *ctx, const smash_text *t, smash_text_info *info);
/* The packed batch a slot holds in HBM, for checks: sizes, then a copy into caller buffers of those sizes
 * (offset arrays n_reads+1; any pointer may be NULL). */
int smash_batch_sizes(smash_ctx *ctx, int slot, uint64_t *n_reads, uint64_t *name_bytes, uint64_t *seq_bytes, uint64_t *opt_bytes);
int smash_fetch_batch(smash_ctx *ctx, int slot, uint8_t *names, int64_t *name_off, uint8_t *seq, uint8_t *qual, int64_t *seq_off,
                      uint8_t *opt, int64_t *opt_off, uint16_t *read_flag);

/* ---- tail: smashMEM.py filter (smashMEM.py:84-92,193-228 with "0 0 10000 4") + awk/perl chromosome
 * filter (smash_mapping.sh:29) + varbin.py (varbin.py:6-118).  Bins are bins.txt column 3
 * (start_abspos, ascending); chrom_* describe chrom_sizes.txt (names matched against the index's
 * forward sequence names).  Counting is done when smash_tail_finish is called. */
int smash_tail_configure(smash_ctx *ctx, const int64_t *bin_starts, uint64_t n_bins,
                         const char *const *chrom_names, const int64_t *chrom_offsets,
                         uint64_t n_chroms, int64_t hit_window, int32_t min_excess);
typedef struct {
  uint64_t total_reads, dups_removed, reads_kept;   /* varbin stats line (varbin.py:104-114) */
  uint64_t n_dupe_pairs, n_non_dupe_pairs;          /* smashMEM.py:230 trailer */
  uint64_t n_positions;                             /* lines of <id>.positions.txt */
} smash_tail_stats;
/* Pair order: smashMEM.py reads a BAM that `samtools sort -n` has put in read-name order, so its first-wins duplicate
 * rule, the order of positions.txt and varbin's adjacent-duplicate rule follow NAME order (samtools 0.1.x strnum_cmp),
 * not arrival order.  The library keeps the name of every pair (reads 2k and 2k+1 are mates and share it), checks the
 * order while batches are appended and, if any pair arrived out of order, sorts the pairs by name on the device before
 * counting -- results do not depend on the order or batching of the input.  The read-sharded phases below need
 * name-ordered input (shards are contiguous ranges) and fail with SMASH_ERR_DATA otherwise.
 * counts: n_bins int64 (this rank's counts; the caller allreduces across GPUs).  If
 * counts_device != NULL the counts are also left in that device buffer (for NCCL). */
int smash_tail_finish(smash_ctx *ctx, int64_t *counts, void *counts_device, smash_tail_stats *st);
/* ---- read-sharded multi-GPU tail (one context per rank; the host moves the small arrays with
 * NCCL/gloo, see smash_paper_b200/multigpu.py).  Ranks hold contiguous ranges of pairs; rank r's pair
 * i has the global ordinal ordinal_base + i.
 *  1. smash_tail_export_keys: device array of {fp1, fp2, ordinal} (3 x u64) for this rank's pairs that
 *     reach smashMEM's dupe set, sorted by ordinal.  The host all-gathers them.
 *  2. smash_tail_phase_a with the concatenated keys of the LOWER ranks (device pointer): global
 *     first-wins duplicate removal (smashMEM.py:217-228) + ordered compaction; returns the shard edge.
 *  3. the host all-gathers the edges; smash_tail_phase_b gets the last filtered position of the
 *     nearest lower rank that has one (varbin.py:56-58 compares with the previous kept line).
 *  4. the host all-reduces the counts and the stats.
 * smash_tail_finish == phase_a(no foreign keys) + phase_b(no predecessor). */
typedef struct { uint64_t n_filtered; int64_t first_pos, last_pos; } smash_tail_edge;
int smash_tail_export_keys(smash_ctx *ctx, uint64_t ordinal_base, const void **dev_keys, uint64_t *n_keys);
int smash_tail_phase_a(smash_ctx *ctx, uint64_t ordinal_base, const void *foreign_keys_dev, uint64_t n_foreign,
                       smash_tail_edge *edge);
/* Variant of phase A for many ranks: the host has already resolved the first-wins rule with a
 * hash-partitioned exchange (every rank owns 1/N of the key space, multigpu.py) and passes, for each of
 * this rank's exported keys (same order as smash_tail_export_keys), the smallest global ordinal that
 * carries that key; a pair survives iff that ordinal is its own. */
int smash_tail_phase_a_verdict(smash_ctx *ctx, uint64_t ordinal_base, const void *min_ordinal_dev, uint64_t n_keys,
                               smash_tail_edge *edge);
int smash_tail_phase_b(smash_ctx *ctx, int has_prev, int64_t prev_last_pos, int64_t *counts, void *counts_device,
                       smash_tail_stats *st);
/* ---- the same protocol behind ONE call, with NCCL moving the data (csrc/comm.cu): this is the "smash_bins_finish
 * (includes the allreduce)" of the survey's boundary.  One communicator rank per context.
 *  - one process per GPU: rank 0 calls smash_comm_unique_id, the host program distributes the 128 bytes (MPI, torchrun's
 *    store, a file ..), every rank calls smash_comm_init_rank;
 *  - one process driving n GPUs (bin/mummer -gpus n): smash_comm_init_all over its n contexts, then one host thread per
 *    context.
 * smash_bins_finish is collective: every rank calls it after collecting its batches.  Rank r must have mapped a
 * contiguous range of the name-ordered read pairs and passes an ordinal_base >= the previous rank's base + pair count
 * (r * 2^40 will do).  The first-wins duplicate rule is resolved on the device by a hash-partitioned exchange of the
 * dupe-set fingerprints (two all-to-alls), the adjacent-duplicate rule by an all-gather of the shard edges, and the
 * n_bins int64 counts (+ stats) are summed by a single ncclAllReduce: every rank receives the GLOBAL counts and stats.
 * Without a communicator it is smash_tail_finish.  NCCL is bound at run time (libnccl.so.2, or SMASH_NCCL_LIB). */
int smash_comm_unique_id(void *id128, size_t cap);
int smash_comm_init_rank(smash_ctx *ctx, int rank, int world, const void *id128);
int smash_comm_init_all(smash_ctx *const *ctxs, int n);
int smash_comm_destroy(smash_ctx *ctx);
int smash_comm_rank(const smash_ctx *ctx, int *rank, int *world);
int smash_bins_finish(smash_ctx *ctx, uint64_t ordinal_base, int64_t *counts, void *counts_device, smash_tail_stats *st);
/* positions.txt rows produced so far by smash_tail_finish: chromosome index (into the forward
 * sequences) and 0-based position, in output order. */
int smash_tail_positions(smash_ctx *ctx, const int32_t **chrom, const int64_t **pos, uint64_t *n);
/* Capacity hint (optional): pre-allocate the tail's HBM buffers for a run of max_pairs read pairs
 * producing at most max_hits kept hits, so no allocation happens inside the run. */
int smash_tail_reserve(smash_ctx *ctx, uint64_t max_pairs, uint64_t max_hits);
/* Forget everything accumulated so far (buffers are kept). */
int smash_tail_reset(smash_ctx *ctx);

/* ---- GC normalisation of the bin counts: the head of cbs.segment01 (cbs.r:18-25) with lowess.gc (cbs.r:3-7), i.e. what
 * binning.sh:39 computes first from varbin's counts and gc.txt:
 *     a <- bincount + 1;  ratio <- a / mean(a[autosomes]);
 *     lowratio <- exp(log(ratio) - approx(lowess(gc.content, log(ratio), f), xout = gc.content)$y)
 * R's stats::lowess (clowess, iter robustness iterations, delta = 1 % of the gc range) and stats::approx (ties averaged).
 * smash_gcnorm_create works out, on the host, everything that depends on gc.content alone (the sort, which points get a
 * local fit and their windows, the tie groups); smash_gcnorm_run does the arithmetic on the GPU for one vector of counts
 * (host pointer, or a device pointer such as the one smash_tail_finish / smash_bins_finish filled).  cbs.r uses f = 0.05
 * and lowess's default iter = 3; autosome[b] != 0 for the bins of chr1..chr22 (cbs.r:13-16, 21).  ratio / lowratio: host
 * arrays of n_bins doubles (either may be null).  Double precision throughout; results agree with R to rounding. */
typedef struct smash_gcnorm smash_gcnorm;
int smash_gcnorm_create(int device, const double *gc_content, const uint8_t *autosome, uint64_t n_bins, double f, int iter,
                        smash_gcnorm **out);
int smash_gcnorm_run(smash_gcnorm *g, const int64_t *counts, const void *counts_device, double *ratio, double *lowratio);
void smash_gcnorm_destroy(smash_gcnorm *g);

/* cudaMemcpy(cudaMemcpyDefault): lets a host language move the small exchange arrays of the multi-GPU
 * tail between library-owned device memory and its own (e.g. torch) tensors. */
int smash_memcpy(void *dst, const void *src, size_t bytes);

/* smash_submit cuts a batch into up to `max_chunks` (1..8) read ranges that flow through separate
 * upload / kernel / download streams, so the SAM text of the first range crosses PCIe while later
 * ranges are still searched; ranges hold at least `min_reads` reads (smaller batches go through whole).
 * `max_chunks` 1..8; defaults 8 / 65536.
 * The output is byte-identical either way (MEM mode and SMASH_WANT_MATCHES always go through whole). */
int smash_ctx_set_chunking(smash_ctx *ctx, int max_chunks, uint64_t min_reads);

/* transport 0 (default): each read range of a batch travels whichever way gets its lines into host memory first --
 * whole text over the download stream, or compact + host threads -- from the measured rate and the backlog of both
 * (PCIe-bound hosts use both at once, hosts short of cores or memory bandwidth fall back to the download stream);
 * 1: always the whole SAM text (A/B runs, tests); 2: always compact.  host_threads: size of the context's
 * line-building pool (0 = one per available core, at most 16; the reference's -qthreads).  Not while batches are in
 * flight.  The output bytes are the same in every case. */
int smash_ctx_set_transport(smash_ctx *ctx, int transport, int host_threads);
/* Bytes copied host->device / device->host by smash_submit* / smash_batch_upload / smash_map_* since the last reset,
 * counted from the copies the library enqueues. */
void smash_ctx_io_bytes(smash_ctx *ctx, uint64_t *h2d, uint64_t *d2h, int reset);
/* Host half of the compact transport, exported for tests (no GPU needed): n_records CmpMeta entries (csrc/compact.h:
 * {u64 sam_off; u32 cmp_off, read, head_len, tags_len, lr_len | rc << 31, pad}) + compact text -> SAM lines. */
int smash_host_expand(const smash_batch *b, uint64_t read_base, const void *meta, uint64_t n_records, const char *cmp, char *sam);

/* ---- counters for bench.py: kernels launched by this library since ctx creation */
uint64_t smash_ctx_launch_count(const smash_ctx *ctx);
/* Device milliseconds accumulated per stage since the last reset, measured with CUDA events on the
 * launching stream: [0] search, [1] records, [2] sizes+scan, [3] emit_text, [4] match CSR, [5] tail, [6] emit_copy,
 * [7] verify (k_mam_verify of the split search, 0 when candidates are verified inside k_mam_search). */
void smash_ctx_stage_ms(smash_ctx *ctx, double *out8, int reset);
/* Device milliseconds of the input stage (smash_submit_text / smash_text_upload) since the last reset: from the end of
 * the text's H2D copy to the end of the kernel that writes the packed batch, CUDA events on the slot's stream. */
double smash_ctx_ingest_ms(smash_ctx *ctx, int reset);
/* Bytes of HBM held by the index on this ctx (text, SA, LCP, seed table, ...). */
uint64_t smash_ctx_index_bytes(const smash_ctx *ctx);
/* Raw CUDA stream (cudaStream_t) the ctx launches on, for external event timing. */
void *smash_ctx_stream(const smash_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SMASH_B200_H_ */
