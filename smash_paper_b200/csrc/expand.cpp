// expand.cpp -- see expand.h.  The work is byte moving: per record ~480 bytes out, of which ~300 are the read's own
// SEQ/QUAL.  Copies are done with overlapping 32/16/8-byte moves (no libc call per field); the reverse complement is a
// byte reversal + two 16-entry table look-ups per 32 bytes (AVX2), with a scalar twin for CPUs without it.
#include "expand.h"

#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace smash {

// reverse_complement's character map (fasta.cpp:26-61): acgt, the IUPAC pairs r/y, m/k, b/v, d/h, either case;
// everything else maps to itself.  Letters are 0x40 | case 0x20 | index 1..26: the map works on the index.
static inline uint8_t comp5(uint8_t i) {
  switch (i) {
    case 1: return 20; case 20: return 1;    // a t
    case 3: return 7; case 7: return 3;      // c g
    case 18: return 25; case 25: return 18;  // r y
    case 13: return 11; case 11: return 13;  // m k
    case 2: return 22; case 22: return 2;    // b v
    case 4: return 8; case 8: return 4;      // d h
    default: return i;
  }
}
static inline uint8_t comp_byte(uint8_t c) { return (c & 0xC0) == 0x40 ? (uint8_t)((c & 0xE0) | comp5(c & 0x1F)) : c; }

static inline void copy_bytes(char *d, const void *s_, size_t n) {
  const char *s = (const char *)s_;
#if defined(__x86_64__)
  if (n >= 16) {
    size_t i = 0;
    for (; i + 16 <= n; i += 16) _mm_storeu_si128((__m128i *)(d + i), _mm_loadu_si128((const __m128i *)(s + i)));
    if (i < n) _mm_storeu_si128((__m128i *)(d + n - 16), _mm_loadu_si128((const __m128i *)(s + n - 16)));
    return;
  }
#endif
  if (n >= 8) { uint64_t a, b; memcpy(&a, s, 8); memcpy(&b, s + n - 8, 8); memcpy(d, &a, 8); memcpy(d + n - 8, &b, 8); return; }
  if (n >= 4) { uint32_t a, b; memcpy(&a, s, 4); memcpy(&b, s + n - 4, 4); memcpy(d, &a, 4); memcpy(d + n - 4, &b, 4); return; }
  for (size_t i = 0; i < n; ++i) d[i] = s[i];
}

static void reverse_complement_scalar(char *dst, const uint8_t *src, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = (char)comp_byte(src[n - 1 - i]);
}
static void reverse_bytes_scalar(char *dst, const uint8_t *src, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = (char)src[n - 1 - i];
}

#if defined(__x86_64__)
struct CompTables {
  alignas(32) uint8_t lo[32], hi[32];
  CompTables() { for (int i = 0; i < 16; ++i) { lo[i] = lo[i + 16] = comp5((uint8_t)i); hi[i] = hi[i + 16] = comp5((uint8_t)(16 + i)); } }
};
static const CompTables g_comp;
__attribute__((target("avx2"))) static void reverse_complement_avx2(char *dst, const uint8_t *src, size_t n) {
  const __m256i TL = _mm256_load_si256((const __m256i *)g_comp.lo), TH = _mm256_load_si256((const __m256i *)g_comp.hi);
  const __m256i REV = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  const __m256i M0F = _mm256_set1_epi8(0x0F), M10 = _mm256_set1_epi8(0x10), ME0 = _mm256_set1_epi8((char)0xE0),
                MC0 = _mm256_set1_epi8((char)0xC0), M40 = _mm256_set1_epi8(0x40);
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    __m256i x = _mm256_loadu_si256((const __m256i *)(src + n - 32 - i));
    x = _mm256_shuffle_epi8(x, REV);                                               // bytes reversed inside each half
    x = _mm256_permute2x128_si256(x, x, 0x01);                                     // halves swapped
    const __m256i lo4 = _mm256_and_si256(x, M0F);
    const __m256i a = _mm256_shuffle_epi8(TL, lo4), b = _mm256_shuffle_epi8(TH, lo4);
    const __m256i hi_sel = _mm256_cmpeq_epi8(_mm256_and_si256(x, M10), M10);
    const __m256i idx = _mm256_blendv_epi8(a, b, hi_sel);
    const __m256i mapped = _mm256_or_si256(_mm256_and_si256(x, ME0), idx);
    const __m256i is_letter = _mm256_cmpeq_epi8(_mm256_and_si256(x, MC0), M40);
    _mm256_storeu_si256((__m256i *)(dst + i), _mm256_blendv_epi8(x, mapped, is_letter));
  }
  for (; i < n; ++i) dst[i] = (char)comp_byte(src[n - 1 - i]);
}
__attribute__((target("avx2"))) static void reverse_bytes_avx2(char *dst, const uint8_t *src, size_t n) {
  const __m256i REV = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    __m256i x = _mm256_loadu_si256((const __m256i *)(src + n - 32 - i));
    x = _mm256_shuffle_epi8(x, REV);
    x = _mm256_permute2x128_si256(x, x, 0x01);
    _mm256_storeu_si256((__m256i *)(dst + i), x);
  }
  for (; i < n; ++i) dst[i] = (char)src[n - 1 - i];
}
static const bool g_avx2 = __builtin_cpu_supports("avx2");
#else
static const bool g_avx2 = false;
#endif

void reverse_complement(char *dst, const uint8_t *src, size_t n) {
#if defined(__x86_64__)
  if (g_avx2) { reverse_complement_avx2(dst, src, n); return; }
#endif
  reverse_complement_scalar(dst, src, n);
}
void reverse_bytes(char *dst, const uint8_t *src, size_t n) {
#if defined(__x86_64__)
  if (g_avx2) { reverse_bytes_avx2(dst, src, n); return; }
#endif
  reverse_bytes_scalar(dst, src, n);
}

// One record's line at d (scalar / SSE moves); returns the line's length.
static inline size_t put_line(const ExpandArgs &a, const CmpMeta &m, char *d0) {
  char *d = d0;
  const uint64_t read = a.read_base + m.read;
  const char *c = a.cmp + m.cmp_off;
  const int64_t no = a.name_off[read];
  const size_t nl = (size_t)(a.name_off[read + 1] - no);
  copy_bytes(d, a.names + no, nl); d += nl;
  copy_bytes(d, c, m.head_len); d += m.head_len; c += m.head_len;
  const int64_t so = a.seq_off[read];
  const size_t q = (size_t)(a.seq_off[read + 1] - so);
  if (m.lr_len & 0x80000000u) {
    reverse_complement(d, a.seq + so, q);
    d[q] = '\t';
    reverse_bytes(d + q + 1, a.qual + so, q);
  } else {
    copy_bytes(d, a.seq + so, q);
    d[q] = '\t';
    copy_bytes(d + q + 1, a.qual + so, q);
  }
  d += 2 * q + 1;
  copy_bytes(d, c, m.tags_len); d += m.tags_len; c += m.tags_len;
  if (a.opt) {
    const int64_t oo = a.opt_off[read];
    const size_t ol = (size_t)(a.opt_off[read + 1] - oo);
    copy_bytes(d, a.opt + oo, ol); d += ol;
  }
  const size_t ll = m.lr_len & 0x7fffffffu;
  copy_bytes(d, c, ll); d += ll;
  return (size_t)(d - d0);
}
static inline size_t line_len(const ExpandArgs &a, const CmpMeta &m) {
  const uint64_t read = a.read_base + m.read;
  size_t n = (size_t)(a.name_off[read + 1] - a.name_off[read]) + m.head_len + 2 * (size_t)(a.seq_off[read + 1] - a.seq_off[read]) + 1 +
             m.tags_len + (m.lr_len & 0x7fffffffu);
  if (a.opt) n += (size_t)(a.opt_off[read + 1] - a.opt_off[read]);
  return n;
}

#if defined(__x86_64__)
// the same with 32-byte moves (the staging buffer has slack after the line, the sources are read with an overlapping tail)
__attribute__((target("avx2"))) static inline void copy32(char *d, const void *s_, size_t n) {
  const char *s = (const char *)s_;
  if (n >= 32) {
    size_t i = 0;
    for (; i + 32 <= n; i += 32) _mm256_storeu_si256((__m256i *)(d + i), _mm256_loadu_si256((const __m256i *)(s + i)));
    if (i < n) _mm256_storeu_si256((__m256i *)(d + n - 32), _mm256_loadu_si256((const __m256i *)(s + n - 32)));
    return;
  }
  copy_bytes(d, s, n);
}
__attribute__((target("avx2"))) static inline size_t put_line_avx2(const ExpandArgs &a, const CmpMeta &m, char *d0) {
  char *d = d0;
  const uint64_t read = a.read_base + m.read;
  const char *c = a.cmp + m.cmp_off;
  const int64_t no = a.name_off[read];
  const size_t nl = (size_t)(a.name_off[read + 1] - no);
  copy32(d, a.names + no, nl); d += nl;
  copy32(d, c, m.head_len); d += m.head_len; c += m.head_len;
  const int64_t so = a.seq_off[read];
  const size_t q = (size_t)(a.seq_off[read + 1] - so);
  if (m.lr_len & 0x80000000u) {
    reverse_complement_avx2(d, a.seq + so, q);
    d[q] = '\t';
    reverse_bytes_avx2(d + q + 1, a.qual + so, q);
  } else {
    copy32(d, a.seq + so, q);
    d[q] = '\t';
    copy32(d + q + 1, a.qual + so, q);
  }
  d += 2 * q + 1;
  copy32(d, c, m.tags_len); d += m.tags_len; c += m.tags_len;
  if (a.opt) {
    const int64_t oo = a.opt_off[read];
    const size_t ol = (size_t)(a.opt_off[read + 1] - oo);
    copy32(d, a.opt + oo, ol); d += ol;
  }
  const size_t ll = m.lr_len & 0x7fffffffu;
  copy32(d, c, ll); d += ll;
  return (size_t)(d - d0);
}

// Lines of consecutive records are consecutive in the SAM text (input order), so a task's output is one long run.
// The run is assembled in a cache-resident staging buffer and leaves as aligned 64-byte NON-TEMPORAL stores: the
// destination lines are never read into the cache first (no read-for-ownership), which is what bounds a plain copy
// when every core of the host writes at once.  Bytes before the first / after the last full cache line of a run (lines
// shared with a neighbouring task) are written with ordinary stores.
struct alignas(64) StreamWriter {
  static constexpr size_t STAGE = 16384, MAX_LINE = 8192;
  char buf[STAGE + MAX_LINE + 64];
  char *dst = nullptr;                                       // destination of buf[fill]
  size_t fill = 0;
};
__attribute__((target("avx2"))) static void sw_flush(StreamWriter &w, bool all) {
  char *base = w.dst - w.fill;                               // destination of buf[0]
  size_t done = 0;
  if (w.fill >= 128) {
    const size_t head = (64 - ((uintptr_t)base & 63)) & 63;
    memcpy(base, w.buf, head);
    const size_t n64 = (w.fill - head) / 64;
    const char *src = w.buf + head;
    char *d = base + head;
    for (size_t i = 0; i < n64; ++i) {
      const __m256i x = _mm256_loadu_si256((const __m256i *)(src + 64 * i)), y = _mm256_loadu_si256((const __m256i *)(src + 64 * i + 32));
      _mm256_stream_si256((__m256i *)(d + 64 * i), x);
      _mm256_stream_si256((__m256i *)(d + 64 * i + 32), y);
    }
    done = head + 64 * n64;
  }
  const size_t rest = w.fill - done;
  if (all) { memcpy(base + done, w.buf + done, rest); w.fill = 0; }
  else { memmove(w.buf, w.buf + done, rest); w.fill = rest; }
}
__attribute__((target("avx2"))) static void expand_records_stream(const ExpandArgs &a, uint64_t f0, uint64_t f1) {
  StreamWriter w;
  for (uint64_t f = f0; f < f1; ++f) {
    const CmpMeta m = a.meta[f];
    char *d = a.sam + m.sam_off;
    if (d != w.dst) {                                        // a new run (first record of the task, sorted output)
      if (w.fill) sw_flush(w, true);
      w.dst = d;
    }
    if (line_len(a, m) > StreamWriter::MAX_LINE) {           // oversized line (long read): straight to its place
      if (w.fill) sw_flush(w, true);
      put_line(a, m, d); w.dst = nullptr;
      continue;
    }
    const size_t n = put_line_avx2(a, m, w.buf + w.fill);
    w.fill += n; w.dst += n;
    if (w.fill >= StreamWriter::STAGE) sw_flush(w, false);
  }
  if (w.fill) sw_flush(w, true);
  _mm_sfence();
}
#endif

void expand_records(const ExpandArgs &a, uint64_t f0, uint64_t f1) {
#if defined(__x86_64__)
  static const bool no_stream = getenv("SMASH_NO_STREAM_STORES") != nullptr;       // A/B switch
  if (g_avx2 && !no_stream) { expand_records_stream(a, f0, f1); return; }
#endif
  for (uint64_t f = f0; f < f1; ++f) put_line(a, a.meta[f], a.sam + a.meta[f].sam_off);
}

}  // namespace smash
