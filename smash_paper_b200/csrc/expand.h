// expand.h -- host half of the compact SAM transport (compact.h): CmpMeta + compact text + the caller's own batch
// -> the SAM lines print_matches (query.cpp:331-415) writes.  Plain C++ (no CUDA), called from the library's host
// worker threads.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "compact.h"

namespace smash {

struct ExpandArgs {
  // the submitted batch (smash_batch; host memory that stays valid until smash_wait)
  const uint8_t *names; const int64_t *name_off;
  const uint8_t *seq; const uint8_t *qual; const int64_t *seq_off;
  const uint8_t *opt; const int64_t *opt_off;            // may be null
  // one read range of it, as downloaded
  uint64_t read_base;                                      // first read of the range
  const CmpMeta *meta; const char *cmp;
  char *sam;                                               // the batch's SAM text
};
// lines of the range's records [f0, f1)
void expand_records(const ExpandArgs &a, uint64_t f0, uint64_t f1);
// reverse_complement (fasta.cpp:26-61) of n bytes; dst and src must not overlap
void reverse_complement(char *dst, const uint8_t *src, size_t n);
void reverse_bytes(char *dst, const uint8_t *src, size_t n);

}  // namespace smash
