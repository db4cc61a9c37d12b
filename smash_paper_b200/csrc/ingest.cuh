// ingest.cuh -- input side on the device: raw SAM text, or the raw text of a FASTQ
// pair, becomes the packed read batch (BatchDev) in HBM without a host-side parser.
//
// Restated semantics (reference file:line):
//  * QueryReader::run, SAM branch (query.cpp:625-648): getline; empty lines are skipped; the line is
//    tokenised by `istringstream >>` (any C-locale whitespace separates fields): name, flag (unsigned;
//    digits only -- what follows the digits inside the same token becomes the NEXT token, as num_get leaves
//    it in the stream), seven ignored fields, seq, errors, then every further token as "\t" + token
//    (add_optional, query.cpp:150-153).  `flag & 64` appends ":0", else `flag & 128` appends ":1".
//  * Aligner::reset (query.cpp:185-201): a trailing ":0" / ":1" is cut off the name and becomes
//    read_flag 65 / 129.
//  * fastqs_to_sam main loop (fastqs_to_sam.cpp:47-95): alternately one record from each file; `in >> ch`
//    skips whitespace INCLUDING blank lines before the '@'/'>' and before the '+'; name = first token of
//    the rest of the header line, second token = optional -> "XO:Z:<optional>"; bases = next line verbatim;
//    '@' records: '+' line, then the error line verbatim; '>' records: errors = bases; N -> Z in bases when a
//    third argument is given; records with an empty bases line print nothing; flags 77 / 141.  The loop ends
//    at the first file that has no further record.
// Deliberate deviation: where the reference's stream state makes it silently reuse the previous line's
// fields (fewer than 11 fields, non-numeric flag, truncated FASTQ record, bases or errors holding interior
// whitespace) or print SEQ and QUAL of different lengths, the ingest reports SMASH_ERR_DATA instead.
//
// Everything here is __host__ __device__: tests/emul runs the same functions on the host.
#pragma once
#include "core.cuh"

namespace smash {

enum IngErr : uint32_t {
  ING_OK = 0,
  ING_FEW_FIELDS = 1,     // SAM line with fewer than 11 fields
  ING_BAD_FLAG = 2,       // flag field does not start with an unsigned integer that fits 32 bits
  ING_LEN_MISMATCH = 3,   // SEQ and QUAL lengths differ
  ING_FQ_AT = 4,          // "Fastq @ parse error" (fastqs_to_sam.cpp:73-75)
  ING_FQ_PLUS = 5,        // "Fastq + parse error" (fastqs_to_sam.cpp:70-72)
  ING_FQ_NAME = 6,        // "Problem reading read name" (fastqs_to_sam.cpp:56-57)
  ING_FQ_TRUNC = 7,       // record cut short by the end of the file
  ING_FQ_COLUMNS = 8,     // bases / errors line is not exactly one token: the SAM columns would shift
  ING_TOO_LONG = 9        // a field longer than 2^31 bytes
};

enum { ING_EMIT = 1, ING_OPT_SAM = 2, ING_OPT_XO = 4, ING_N2Z = 8 };

// One input line (SAM) or one FASTQ record, located in its source text.
struct LineRec {
  uint64_t name_pos, seq_pos, qual_pos, opt_pos;   // byte offsets into text[src]
  uint32_t name_len, seq_len;
  uint32_t opt_len;        // bytes this read adds to the batch's `opt` blob
  uint32_t opt_src_len;    // ING_OPT_SAM: source bytes from opt_pos-1 (the separator) to the end of the line;
                           // ING_OPT_XO: length of the header's second token
  uint16_t read_flag;      // 0 / 65 / 129
  uint8_t src;             // which text (0: SAM text or FASTQ mate 1, 1: FASTQ mate 2)
  uint8_t bits;            // ING_EMIT | ING_OPT_* | ING_N2Z
  uint32_t err;            // IngErr
};

// what the scan adds up per record: reads emitted, name / seq / opt bytes
struct Ing4 { uint64_t reads, name, seq, opt; };
HD Ing4 ing4_add(const Ing4 &a, const Ing4 &b) { return Ing4{a.reads + b.reads, a.name + b.name, a.seq + b.seq, a.opt + b.opt}; }
HD Ing4 ing4_of(const LineRec &r) {
  if (!(r.bits & ING_EMIT)) return Ing4{0, 0, 0, 0};
  return Ing4{1, r.name_len, r.seq_len, r.opt_len};
}

// C-locale isspace: what `operator>>` skips and what ends a token
HD bool ing_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// ---- SWAR over 8 text bytes (little-endian word) ------------------------------------------------------
constexpr uint64_t ING_LO7 = 0x7f7f7f7f7f7f7f7full, ING_HI = 0x8080808080808080ull, ING_ONES = 0x0101010101010101ull;
HD int ing_ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
HD int ing_popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
// 0x80 in every byte of w that equals c (exact: no carries between bytes)
HD uint64_t ing_eq_mask(uint64_t w, uint8_t c) {
  const uint64_t x = w ^ (ING_ONES * c);
  const uint64_t t = ((x & ING_LO7) + ING_LO7) | x;
  return ~t & ING_HI;
}
// 0x80 in every whitespace byte (0x20, 0x09..0x0d)
HD uint64_t ing_ws_mask(uint64_t w) {
  const uint64_t w7 = (w & ING_LO7) | ING_HI;
  const uint64_t ge9 = w7 - ING_ONES * 9, ge14 = w7 - ING_ONES * 14;     // top bit of a byte stays set iff (byte & 0x7f) >= 9 / 14
  return (ge9 & ~ge14 & ~w & ING_HI) | ing_eq_mask(w, ' ');
}
// 8 bytes at any offset, from the two aligned words that hold them.  May touch up to 15 bytes past t[i+7]: every
// text buffer carries 64 bytes of slack (device: ing_raw; host emulation: padded copies).
HD uint64_t ing_load8(const uint8_t *t, uint64_t i) {
  const uintptr_t a = (uintptr_t)(t + i);
  const uint64_t *p = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
  const unsigned sh = (unsigned)(a & 7) * 8u;
  const uint64_t lo = p[0];
  if (!sh) return lo;
  return (lo >> sh) | (p[1] << (64u - sh));
}
// end of the token that starts at or after i (first whitespace byte, or e): 8 bytes per step, every aligned word of
// the text loaded once
HD uint64_t ing_token_end(const uint8_t *t, uint64_t i, uint64_t e) {
  if (i + 8 <= e) {
    const uintptr_t a = (uintptr_t)(t + i);
    const uint64_t *p = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(a & 7) * 8u;
    uint64_t lo = *p;
    do {
      uint64_t w = lo;
      if (sh) { const uint64_t hi = *++p; w = (lo >> sh) | (hi << (64u - sh)); lo = hi; } else { lo = *++p; }
      const uint64_t m = ing_ws_mask(w);
      if (m) return i + (uint64_t)(ing_ctz64(m) >> 3);
      i += 8;
    } while (i + 8 <= e);
  }
  while (i < e && !ing_space(t[i])) ++i;
  return i;
}

// ---- lines ------------------------------------------------------------------------------------
// A line starts at 0 and after every '\n' that is not the last byte.  Line starts are counted per 16-byte
// chunk of the text (two little-endian words lo, hi); a start is attributed to the chunk that holds the
// newline BEFORE it (the first line to chunk 0), which keeps them in order.
HD uint64_t ing_low_bytes(int k) { return k >= 8 ? ~0ull : (k <= 0 ? 0ull : ((1ull << (8 * k)) - 1ull)); }
HD void ing_chunk_masks(uint64_t lo, uint64_t hi, uint64_t base, uint64_t n, uint64_t *m0, uint64_t *m1) {
  *m0 = ing_eq_mask(lo, '\n'); *m1 = ing_eq_mask(hi, '\n');
  if (base + 17 > n) {                                   // last chunk: only newlines at j with j + 1 < n start a line
    const int keep = (int)(n - 1 - base);                // bytes 0 .. keep-1 of the chunk qualify (n > base)
    *m0 &= ing_low_bytes(keep); *m1 &= ing_low_bytes(keep - 8);
  }
}
HD uint32_t ing_chunk_starts(uint64_t lo, uint64_t hi, uint64_t base, uint64_t n) {
  uint64_t m0, m1;
  ing_chunk_masks(lo, hi, base, n, &m0, &m1);
  return ((base == 0 && n > 0) ? 1u : 0u) + (uint32_t)ing_popc64(m0) + (uint32_t)ing_popc64(m1);
}
// writes the chunk's line starts to ls[idx...]; returns the next idx
HD uint64_t ing_chunk_place(uint64_t lo, uint64_t hi, uint64_t base, uint64_t n, uint64_t *ls, uint64_t idx) {
  uint64_t m0, m1;
  ing_chunk_masks(lo, hi, base, n, &m0, &m1);
  if (base == 0 && n > 0) ls[idx++] = 0;
  while (m0) { ls[idx++] = base + (uint64_t)(ing_ctz64(m0) >> 3) + 1; m0 &= m0 - 1; }
  while (m1) { ls[idx++] = base + 8 + (uint64_t)(ing_ctz64(m1) >> 3) + 1; m1 &= m1 - 1; }
  return idx;
}
// content of line j = [ls[j], ing_line_end(ls, j)); ls[n_lines] is a sentinel placed one past the
// (possibly absent) newline of the last line
HD uint64_t ing_line_end(const uint64_t *ls, uint64_t j) { return ls[j + 1] - 1; }
HD uint64_t ing_sentinel(const uint8_t *text, uint64_t n) { return n == 0 ? 1 : (text[n - 1] == '\n' ? n : n + 1); }

// ---- SAM line (query.cpp:639-648 + 185-201) ------------------------------------------------------
HD void ing_name_flag(const uint8_t *t, LineRec &r, uint32_t flag) {
  if (flag & 64) r.read_flag = 65;                       // name + ":0" -> stripped again
  else if (flag & 128) r.read_flag = 129;
  else {
    r.read_flag = 0;
    if (r.name_len >= 2 && t[r.name_pos + r.name_len - 2] == ':') {
      const uint8_t c = t[r.name_pos + r.name_len - 1];
      if (c == '0') { r.name_len -= 2; r.read_flag = 65; }
      else if (c == '1') { r.name_len -= 2; r.read_flag = 129; }
    }
  }
}

HDN inline void ing_parse_sam_line(const uint8_t *t, uint64_t b, uint64_t e, LineRec &r) {
  r = LineRec{};
  if (e <= b) return;                                     // empty line: skipped (query.cpp:627)
  r.bits = ING_EMIT;
  uint64_t i = b;
#define ING_SKIP_WS() while (i < e && ing_space(t[i])) ++i
#define ING_TOKEN() i = ing_token_end(t, i, e)
  ING_SKIP_WS();
  if (i == e) { r.err = ING_FEW_FIELDS; return; }
  r.name_pos = i;
  ING_TOKEN();
  if (i - r.name_pos > 0x7fffffffull) { r.err = ING_TOO_LONG; return; }
  r.name_len = (uint32_t)(i - r.name_pos);
  // flag: [+-]digits (num_get for unsigned, base 10); stops at the first non-digit
  ING_SKIP_WS();
  if (i == e) { r.err = ING_FEW_FIELDS; return; }
  bool neg = false;
  if (t[i] == '+' || t[i] == '-') { neg = t[i] == '-'; ++i; }
  uint64_t v = 0; int digits = 0; bool over = false;
  while (i < e && t[i] >= '0' && t[i] <= '9') {
    v = v * 10 + (uint64_t)(t[i] - '0');
    if (v > 0xffffffffull) { over = true; v = 0xffffffffull; }
    ++digits; ++i;
  }
  if (!digits || over) { r.err = ING_BAD_FLAG; return; }
  const uint32_t flag = neg ? (uint32_t)(0u - (uint32_t)v) : (uint32_t)v;
  // seven ignored fields; whatever followed the digits inside the flag token is the first of them
  for (int k = 0; k < 7; ++k) {
    ING_SKIP_WS();
    if (i == e) { r.err = ING_FEW_FIELDS; return; }
    ING_TOKEN();
  }
  ING_SKIP_WS();
  if (i == e) { r.err = ING_FEW_FIELDS; return; }
  r.seq_pos = i;
  ING_TOKEN();
  const uint64_t sl = i - r.seq_pos;
  ING_SKIP_WS();
  if (i == e) { r.err = ING_FEW_FIELDS; return; }
  r.qual_pos = i;
  ING_TOKEN();
  const uint64_t ql = i - r.qual_pos;
  if (sl > 0x7fffffffull) { r.err = ING_TOO_LONG; return; }
  if (sl != ql) { r.err = ING_LEN_MISMATCH; return; }
  r.seq_len = (uint32_t)sl;
  // optional fields: "\t" + token for every further token
  const uint64_t sep = i;                                  // the separator before the first optional token (or e)
  uint64_t ol = 0;
  for (;;) {
    ING_SKIP_WS();
    if (i == e) break;
    if (!ol) r.opt_pos = i;
    const uint64_t s = i;
    ING_TOKEN();
    ol += 1 + (i - s);
  }
#undef ING_SKIP_WS
#undef ING_TOKEN
  if (ol) {
    if (ol > 0x7fffffffull || e - sep > 0x7fffffffull) { r.err = ING_TOO_LONG; return; }
    r.bits |= ING_OPT_SAM;
    r.opt_len = (uint32_t)ol;
    r.opt_src_len = (uint32_t)(e - (r.opt_pos - 1));      // from the whitespace byte before the first token
  }
  ing_name_flag(t, r, flag);
}

// Optional-field text of a SAM line: source byte i of [opt_pos-1, line end) -> output byte, or -1 for none.
// Token bytes pass through; the LAST whitespace byte before a token becomes the single '\t'.
HD int ing_opt_char(const uint8_t *t, uint64_t i, uint64_t e) {
  const uint8_t c = t[i];
  if (!ing_space(c)) return c;
  return (i + 1 < e && !ing_space(t[i + 1])) ? '\t' : -1;
}

// ---- FASTQ (fastqs_to_sam.cpp:47-95) ---------------------------------------------------------------
// The reader is a 6-state machine over LINES; a line's effect on the state depends only on whether it is
// blank (whitespace only) and on its first non-blank byte, so the state before every line follows from an
// exclusive scan of per-line transition functions under composition.
enum { FQ_H = 0, FQ_BA = 1, FQ_BG = 2, FQ_P = 3, FQ_Q = 4, FQ_E = 5 };   // expect: header, bases ('@'), bases ('>'), plus, errors; error
constexpr uint32_t FQ_IDENT = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12) | (5u << 15);
HD uint32_t fq_pack(int h, int ba, int bg, int p, int q, int e) {
  return (uint32_t)h | ((uint32_t)ba << 3) | ((uint32_t)bg << 6) | ((uint32_t)p << 9) | ((uint32_t)q << 12) | ((uint32_t)e << 15);
}
HD int fq_apply(uint32_t f, int s) { return (int)((f >> (3 * s)) & 7u); }
HD uint32_t fq_compose(uint32_t f, uint32_t g) {             // first f, then g
  uint32_t h = 0;
#pragma unroll
  for (int s = 0; s < 6; ++s) h |= (uint32_t)fq_apply(g, fq_apply(f, s)) << (3 * s);
  return h;
}
// first non-blank byte of the line, or -1 when the line is blank
HD int ing_first_char(const uint8_t *t, uint64_t b, uint64_t e) {
  for (uint64_t i = b; i < e; ++i) if (!ing_space(t[i])) return t[i];
  return -1;
}
HD uint32_t fq_line_fn(int first) {
  //                         H      BA    BG    P     Q     E
  if (first < 0) return fq_pack(FQ_H, FQ_P, FQ_H, FQ_P, FQ_H, FQ_E);            // blank: skipped where `>> ch` reads
  if (first == '@') return fq_pack(FQ_BA, FQ_P, FQ_H, FQ_E, FQ_H, FQ_E);
  if (first == '>') return fq_pack(FQ_BG, FQ_P, FQ_H, FQ_E, FQ_H, FQ_E);
  if (first == '+') return fq_pack(FQ_E, FQ_P, FQ_H, FQ_Q, FQ_H, FQ_E);
  return fq_pack(FQ_E, FQ_P, FQ_H, FQ_E, FQ_H, FQ_E);
}

// exactly one token on [b,e) (leading / trailing whitespace, e.g. '\r', allowed): its position and length
HD bool ing_single_token(const uint8_t *t, uint64_t b, uint64_t e, uint64_t *pos, uint64_t *len) {
  uint64_t i = b;
  while (i < e && ing_space(t[i])) ++i;
  if (i == e) return false;
  *pos = i;
  i = ing_token_end(t, i, e);
  *len = i - *pos;
  while (i < e && ing_space(t[i])) ++i;
  return i == e;
}

// Record whose header is line j of this file (ls: line starts with sentinel, n_lines lines).
HDN inline void ing_parse_fastq_record(const uint8_t *t, const uint64_t *ls, uint64_t n_lines, uint64_t j, int file,
                                       int replace_n, LineRec &r) {
  r = LineRec{};
  r.src = (uint8_t)file;
  r.read_flag = file ? 129 : 65;                           // flags 77 / 141 (fastqs_to_sam.cpp:78), query.cpp:643-644
  uint64_t i = ls[j];
  const uint64_t e = ing_line_end(ls, j);
  while (i < e && ing_space(t[i])) ++i;
  const uint8_t amp = t[i];                                // the caller guarantees a non-blank line
  ++i;
  while (i < e && ing_space(t[i])) ++i;
  if (i == e) { r.err = ING_FQ_NAME; return; }
  r.name_pos = i;
  while (i < e && !ing_space(t[i])) ++i;
  if (i - r.name_pos > 0x7fffffffull) { r.err = ING_TOO_LONG; return; }
  r.name_len = (uint32_t)(i - r.name_pos);
  while (i < e && ing_space(t[i])) ++i;
  if (i < e) {                                             // second token -> "\tXO:Z:" + token
    r.opt_pos = i;
    while (i < e && !ing_space(t[i])) ++i;
    if (i - r.opt_pos > 0x7ffffff0ull) { r.err = ING_TOO_LONG; return; }
    r.opt_src_len = (uint32_t)(i - r.opt_pos);
    r.opt_len = r.opt_src_len + 6;
    r.bits |= ING_OPT_XO;
  }
  if (j + 1 >= n_lines) { r.err = ING_FQ_TRUNC; return; }
  const uint64_t bb = ls[j + 1], be = ing_line_end(ls, j + 1);
  const bool printed = be > bb;                            // `if (bases.size())`, fastqs_to_sam.cpp:76
  uint64_t sp = 0, sl = 0, qp = 0, ql = 0;
  if (printed && !ing_single_token(t, bb, be, &sp, &sl)) { r.err = ING_FQ_COLUMNS; return; }
  if (amp == '@') {
    uint64_t p = j + 2;
    while (p < n_lines && ing_first_char(t, ls[p], ing_line_end(ls, p)) < 0) ++p;
    if (p >= n_lines) { r.err = ING_FQ_TRUNC; return; }
    if (ing_first_char(t, ls[p], ing_line_end(ls, p)) != '+') { r.err = ING_FQ_PLUS; return; }
    if (p + 1 >= n_lines) { r.err = ING_FQ_TRUNC; return; }
    if (printed && !ing_single_token(t, ls[p + 1], ing_line_end(ls, p + 1), &qp, &ql)) { r.err = ING_FQ_COLUMNS; return; }
  } else if (amp == '>') {
    qp = sp; ql = sl;                                      // errors = bases, before the N -> Z replacement
  } else {
    r.err = ING_FQ_AT; return;
  }
  if (!printed) { r.opt_len = 0; r.bits = 0; return; }
  if (sl > 0x7fffffffull) { r.err = ING_TOO_LONG; return; }
  if (sl != ql) { r.err = ING_LEN_MISMATCH; return; }
  r.seq_pos = sp; r.seq_len = (uint32_t)sl; r.qual_pos = qp;
  r.bits |= ING_EMIT;
  if (replace_n) r.bits |= ING_N2Z;
}

// How many records the reference's loop prints from the file it reads FIRST in a round (ra records available)
// and from the other one (rb) (fastqs_to_sam.cpp:47-53): it stops at the first file without a further record.
// final = 0: the last record of either text may be cut by the chunk edge and is left out.
HD void ing_fastq_take(uint64_t ra, uint64_t rb, int final, uint64_t *na, uint64_t *nb) {
  if (!final) { ra = ra ? ra - 1 : 0; rb = rb ? rb - 1 : 0; }
  *na = ra < rb + 1 ? ra : rb + 1;
  *nb = rb < *na ? rb : *na;
}
// Records alternate between the mate files; a chunk of a stream may have to START with mate 2 (phase 1) because the
// previous chunk ended after a mate-1 record.  Record k of file f sits at position 2k + (f ^ phase).
HD uint64_t ing_fastq_pos(uint64_t k, int file, int phase) { return 2 * k + (uint64_t)((file ^ phase) & 1); }
// records of file f among the first m positions
HD uint64_t ing_fastq_taken(uint64_t m, int file, int phase) { return ((file ^ phase) & 1) ? m / 2 : (m + 1) / 2; }

}  // namespace smash
