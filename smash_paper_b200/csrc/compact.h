// compact.h -- the compact SAM transport (shared by the kernels, the host library and the host expander).
//
// A SAM line is  NAME head SEQ \t QUAL tags [optional fields] lr-tags \n  (print_matches, query.cpp:331-415).  NAME, SEQ,
// QUAL and the optional fields are the caller's own input bytes (SEQ/QUAL reversed for reverse-strand records), so only
// the text the GPU actually computes -- head = "\tFLAG\tRNAME\tPOS\tMAPQ\tCIGAR\tRNEXT\tPNEXT\tTLEN\t", the tags, the L/R
// tags and the newline -- crosses PCIe, plus one CmpMeta per record; the host rebuilds the identical line from the
// batch it submitted (expand.cpp).
#pragma once
#include <stdint.h>

namespace smash {

struct CmpMeta {
  uint64_t sam_off;      // where the line starts in the batch's SAM text
  uint32_t cmp_off;      // where head|tags|lr-tags\n start in the range's compact text
  uint32_t read;         // read index inside the range
  uint32_t head_len;     // bytes between NAME and SEQ
  uint32_t tags_len;     // bytes between QUAL and the optional fields
  uint32_t lr_len;       // L/R tags + newline; bit 31: SEQ/QUAL are printed reverse-complemented / reversed
  uint32_t pad;
};
static_assert(sizeof(CmpMeta) == 32, "CmpMeta is two 16-byte words");

}  // namespace smash
