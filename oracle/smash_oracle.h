/* oracle/smash_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference hot path (yamrom/smash-paper: longSA.cpp, query.cpp,
 * memsam.h).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (smash_paper_b200/csrc) never does.
 *
 * Parity status: PINNED.  The restatement is checked byte-for-byte against the unmodified
 * reference compiled into oracle/_ref (see oracle/Makefile, tests/test_oracle_vs_reference.py)
 * and against the committed fixtures in tests/golden/ that the same reference generated
 * (tests/golden/make_golden.py).  The smashMEM.py stage cannot run here (pysam/samtools absent):
 * oracle/tail.py restates it from the source only -> that one stage is "parity unpinned".
 */
#ifndef SMASH_ORACLE_H_
#define SMASH_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t idx;
  uint64_t val; /* low 4 or 8 bytes significant, see orc_index.w */
} orc_lcp_item;   /* on-disk item_t is 16 bytes for both int widths (longSA.h:19-28) */

typedef struct {
  const uint8_t *text;   /* rc{r}.ref.seq.bin, N bytes */
  uint64_t N;
  const void *sa;        /* N * w bytes */
  const void *isa;       /* N * w bytes */
  int w;                 /* sizeof(ANINT): 4 or 8 (size.h:9-22) */
  const uint8_t *lcp_vec;
  const orc_lcp_item *lcp_m;
  uint64_t n_m;
  uint64_t n_descr;      /* 2 * n_chrom with -rcref */
  const uint64_t *startpos;
  const uint64_t *sizes;
  const char *const *descr;
  int rcref;
} orc_index;

typedef struct {
  uint64_t ref, query, len;   /* match_t, longSA.h:78-92 */
} orc_match;

enum { ORC_MUM = 0, ORC_MAM = 1, ORC_MEM = 2 };

/* Index construction for SMALL texts (tests only): plain comparison sort + Kasai.
 * sa/isa are uint64 arrays of N entries, lcp a uint64 array of N entries (full values). */
void orc_build_index(const uint8_t *text, uint64_t N, uint64_t *sa, uint64_t *isa, uint64_t *lcp);

/* longSA::MAM (longSA.cpp:503-536) / longSA::MEM (longSA.cpp:587-590, 395-490).
 * `query` is the already-lowercased read.  Returns the number of matches (all of them are
 * counted even if cap is exceeded; only the first cap are stored). */
uint64_t orc_mam(const orc_index *ix, const uint8_t *query, uint64_t qlen, uint64_t min_len,
                 orc_match *out, uint64_t cap);
uint64_t orc_mem(const orc_index *ix, const uint8_t *query, uint64_t qlen, uint64_t min_len,
                 orc_match *out, uint64_t cap);
/* Independent brute-force statement of SURVEY.md Appendix A.1 (no SA): O(q*N), tiny inputs. */
uint64_t orc_mam_bruteforce(const uint8_t *text, uint64_t N, const uint8_t *query, uint64_t qlen,
                            uint64_t min_len, orc_match *out, uint64_t cap);

/* Whole mummer -samin -samout pass over a packed batch (query.cpp:481-520, 231-434):
 * reads 2k,2k+1 are mates-by-arrival.  read_flag[i] is 0/65/129 (Aligner::reset).  Writes SAM
 * record lines (no header) in input order into out (cap bytes) and returns the byte count
 * needed (call again with a bigger buffer if > cap).  Optionally returns all matches as CSR. */
typedef struct {
  uint64_t n_reads;
  const uint8_t *names;  const int64_t *name_off;
  const uint8_t *seq;    const uint8_t *qual; const int64_t *seq_off;
  const uint8_t *opt;    const int64_t *opt_off;
  const uint16_t *read_flag;
} orc_batch;

typedef struct {
  int mode;              /* ORC_MAM / ORC_MEM / ORC_MUM */
  uint32_t min_len;
  int nomap;
  int nucleotides_only;  /* -n */
  int n_threads;         /* >=1; reads are split in contiguous pair ranges */
} orc_params;

uint64_t orc_map_batch(const orc_index *ix, const orc_batch *b, const orc_params *p,
                       char *out, uint64_t cap,
                       int64_t *match_off /* n_reads+1 or NULL */,
                       orc_match *matches, uint64_t match_cap);

/* longSA::show (longSA.cpp:612-690) with bin=true: map.bin body (2 bytes per forward base, the
 * 2 junk header bytes are NOT produced). out must hold 2*sum(forward sizes) bytes. */
void orc_mappability(const orc_index *ix, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
