// ingest_launch.cuh -- launchers of ingest.cu, called by the host side of the ABI (api.cu).
#pragma once
#include <cuda_runtime.h>

#include "ingest.cuh"

namespace smash {

uint64_t ing_tiles(uint64_t n);            // scan tiles over n items: scratch arrays need ing_tiles(n) + 1 entries

struct IngPublish {
  const Ing4 *pre; uint64_t m;             // exclusive scan over the m records (+ total at [m])
  int final, fastq;
  int phase;                               // FASTQ: 1 = this chunk starts with a mate-2 record
  const uint64_t *ls[2];                   // line starts per text
  const uint64_t *hdr[2];                  // FASTQ: record -> header line
  uint64_t n_rec[2];                       // SAM: n_lines; FASTQ: records per text
  uint64_t n_bytes[2];
  const unsigned long long *err;           // min over (record index << 8 | IngErr), ~0 when clean
  uint64_t *host;                          // mapped pinned: [0..7] reads, name, seq, opt bytes, records used, consumed[2], err; [12] next phase
};
struct IngCopy {
  const uint8_t *text[2];
  const LineRec *recs; const Ing4 *pre; uint64_t m;
  uint8_t *names; int64_t *name_off; uint8_t *seq; uint8_t *qual; int64_t *seq_off;
  uint8_t *opt; int64_t *opt_off;          // null when the batch has no optional fields
  uint16_t *read_flag;
};

// all asynchronous on `st`; each returns the number of kernels it launched
int launch_ing_count_lines(const uint8_t *text, uint64_t n, uint64_t *blk, cudaStream_t st);      // blk[ing_tiles((n+15)/16)] = n_lines
int launch_ing_line_starts(const uint8_t *text, uint64_t n, const uint64_t *blk, uint64_t *ls, cudaStream_t st);
int launch_ing_parse_sam(const uint8_t *text, const uint64_t *ls, uint64_t n_lines, LineRec *recs, unsigned long long *err, cudaStream_t st);
int launch_ing_fastq_headers(const uint8_t *text, const uint64_t *ls, uint64_t n_lines, uint32_t *blk32, uint64_t *blk64, uint8_t *hdr_flag,
                             uint64_t *hdr, uint64_t *count, cudaStream_t st);
int launch_ing_parse_fastq(const uint8_t *text, const uint64_t *ls, uint64_t n_lines, const uint64_t *hdr, uint64_t n_take, int file,
                           int phase, int replace_n, LineRec *recs, unsigned long long *err, cudaStream_t st);
int launch_ing_scan_recs(const LineRec *recs, uint64_t m, Ing4 *blk4, Ing4 *pre, cudaStream_t st);
int launch_ing_publish(const IngPublish &p, cudaStream_t st);
int launch_ing_copy(const IngCopy &c, cudaStream_t st);

}  // namespace smash
