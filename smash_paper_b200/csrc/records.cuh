// records.cuh -- per-read record construction (Alignment::resolve + prepare_matches + set_nomap,
// query.cpp:68-97, 231-320) and SAM line formatting (print_matches, query.cpp:331-415) as
// lane-serial building blocks.  See DESIGN.md "records" / "sam_emit".
#pragma once
#include "core.cuh"

namespace smash {

// One alignment while a read is being processed (scratch, shared memory).
struct Aln {
  int64_t rcpos, pos;
  uint32_t si;
  uint16_t prefix, len, suffix, qpos;      // qpos may be lowered by the merge (query.cpp:283)
  uint16_t xm, xu;                         // n_matches, n_unique_bases (moved along the group)
  uint8_t rc, last;                        // last: this alignment carries its group's record
  uint16_t first_item, last_item;          // group's first/last position in merge order
};

// One printed record in HBM (records of a read are stored in HI order => prev/next implicit).
struct Rec {
  int64_t pos;           // 0-based position on chromosome si
  int64_t rcpos;         // text position of read offset 0 on the matched strand (for XE)
  uint32_t si;           // even seq_index (forward copy)
  uint16_t item_begin, item_cnt;   // CIGAR items (merge order) in the read's item slots
  uint16_t xu, xe, qlen, suffix;   // qlen = read length minus soft clips (pysam qlen)
  uint8_t rc, L0, R0, flags;       // L0/R0: mappability_tag values of the first '=' block
  uint16_t seq_off, lr_len;        // offset of the SEQ column in the line, length of the L/R tags (set by k_sizes)
};
// One '=' block of a record's CIGAR; L/R: its mappability_tag values (set by k_rec_xe when map.bin is loaded, so that the
// size, emit and tail kernels do not go back to map.bin)
struct Item { uint16_t prefix, len; uint8_t L, R; uint16_t pad; };

// Per-read summary (what the mate and the size/emit kernels need).
struct ReadSum {
  int64_t best_pos;
  uint32_t best_si;
  uint16_t n_rec;        // records to print (1 for an unmapped placeholder)
  uint8_t has_best;      // best_alignment != nullptr
  uint8_t unmapped;      // placeholder present (read_flag |= 4)
};

// ---- libstdc++ std::sort on an index array (bits/stl_algo.h), needed because to_print has real
// ties in MEM mode: threshold 16, median-of-3 to first, unguarded
// partition, heap sort when the depth limit is hit, final insertion sort.
template <class Less>
struct StdSort {
  uint16_t *v; Less lt;
  HDN void linear_insert(int last) {
    uint16_t val = v[last]; int next = last - 1;
    while (lt(val, v[next])) { v[last] = v[next]; last = next; --next; }
    v[last] = val;
  }
  HDN void insertion(int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
      if (lt(v[i], v[first])) {
        uint16_t val = v[i];
        for (int j = i; j > first; --j) v[j] = v[j - 1];
        v[first] = val;
      } else linear_insert(i);
    }
  }
  HDN void push_heap(int first, int hole, int top, uint16_t val) {
    int parent = (hole - 1) / 2;
    while (hole > top && lt(v[first + parent], val)) {
      v[first + hole] = v[first + parent]; hole = parent; parent = (hole - 1) / 2;
    }
    v[first + hole] = val;
  }
  HDN void adjust_heap(int first, int hole, int len, uint16_t val) {
    const int top = hole; int child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (lt(v[first + child], v[first + child - 1])) child--;
      v[first + hole] = v[first + child]; hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      v[first + hole] = v[first + child - 1]; hole = child - 1;
    }
    push_heap(first, hole, top, val);
  }
  HDN void heap_sort(int first, int last) {
    int len = last - first;
    if (len >= 2)
      for (int parent = (len - 2) / 2;; --parent) { adjust_heap(first, parent, len, v[first + parent]); if (parent == 0) break; }
    while (last - first > 1) {
      --last; uint16_t val = v[last]; v[last] = v[first];
      adjust_heap(first, 0, last - first, val);
    }
  }
  HDN void swp(int a, int b) { uint16_t t = v[a]; v[a] = v[b]; v[b] = t; }
  HDN void median_to_first(int res, int a, int b, int c) {
    if (lt(v[a], v[b])) {
      if (lt(v[b], v[c])) swp(res, b); else if (lt(v[a], v[c])) swp(res, c); else swp(res, a);
    } else if (lt(v[a], v[c])) swp(res, a);
    else if (lt(v[b], v[c])) swp(res, c);
    else swp(res, b);
  }
  HDN int partition(int first, int last, int pivot) {
    for (;;) {
      while (lt(v[first], v[pivot])) ++first;
      --last;
      while (lt(v[pivot], v[last])) --last;
      if (!(first < last)) return first;
      swp(first, last); ++first;
    }
  }
  HDN void run(int n) {
    if (n <= 0) return;
    if (n > 16) {
      // __introsort_loop with an explicit stack (the recursion is on the right part)
      int lg = 0; for (int k = n; k > 1; k >>= 1) ++lg;
      int stk_first[40], stk_last[40], stk_depth[40], sp = 0;
      stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
      while (sp) {
        --sp; int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
          if (depth == 0) { heap_sort(first, last); break; }
          --depth;
          int mid = first + (last - first) / 2;
          median_to_first(first, first + 1, mid, last - 1);
          int cut = partition(first + 1, last, first);
          if (sp < 40) { stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; ++sp; }
          last = cut;
        }
      }
      insertion(0, 16);
      for (int i = 16; i != n; ++i) linear_insert(i);
    } else insertion(0, n);
  }
};
// NOTE on the explicit stack: libstdc++ recurses on [cut,last) FIRST and then continues with
// [first,cut).  The two halves are disjoint, so processing order does not change the result.

struct LessMerge {   // to_merge, query.cpp:203-219
  const Aln *a;
  HDN bool operator()(uint16_t x, uint16_t y) const {
    const Aln &p = a[x], &q = a[y];
    if (p.rc != q.rc) return p.rc < q.rc;
    if (p.si != q.si) return p.si < q.si;
    if (p.pos != q.pos) return p.pos < q.pos;
    return p.prefix < q.prefix;
  }
};
struct LessPrint {   // to_print, query.cpp:221-229
  const Aln *a;
  HDN bool operator()(uint16_t x, uint16_t y) const {
    const Aln &p = a[x], &q = a[y];
    if (p.qpos == q.qpos) return p.rc < q.rc;
    return p.qpos < q.qpos;
  }
};

struct LessByRef {   // by_ref, longSA.cpp:492-499
  const Match *m;
  HDN bool operator()(uint16_t x, uint16_t y) const {
    if (m[x].ref == m[y].ref) return m[x].len > m[y].len;
    return m[x].ref < m[y].ref;
  }
};
// longSA::MUM (longSA.cpp:549-585) after the MAM search: src[0..n) are the MAM matches in emission
// order; dst receives the survivors of the cleanMUMcand sweep in by_ref order.  Returns their count.
HDN inline int mum_clean(const Match *src, int n, uint16_t *ord, Match *dst) {
  if (n <= 0) return 0;
  for (int i = 0; i < n; ++i) ord[i] = (uint16_t)i;
  StdSort<LessByRef> s{ord, LessByRef{src}}; s.run(n);
  int out = 0; uint64_t dbright = 0; bool ignoreprevious = false;
  for (int i = 0; i < n; ++i) {
    bool ignorecurrent = false;
    const uint64_t currentright = src[ord[i]].ref + src[ord[i]].len - 1;
    if (dbright > currentright) ignorecurrent = true;
    else if (dbright == currentright) {
      ignorecurrent = true;
      if (!ignoreprevious && i > 0 && src[ord[i - 1]].ref == src[ord[i]].ref) ignoreprevious = true;
    } else dbright = currentright;
    if (i > 0 && !ignoreprevious) dst[out++] = src[ord[i - 1]];
    ignoreprevious = ignorecurrent;
  }
  if (!ignoreprevious) dst[out++] = src[ord[n - 1]];
  return out;
}

// Alignment::resolve (query.cpp:68-97)
HD void resolve_match(const DevIndex &ix, const Match &m, int q, Aln *a) {
  int lo = 0, hi = ix.n_descr;                       // upper_bound(startpos, ref)
  while (lo < hi) { int mid = (lo + hi) >> 1; if (ix.startpos[mid] <= m.ref) lo = mid + 1; else hi = mid; }
  uint32_t si = (uint32_t)(lo - 1);
  a->rcpos = (int64_t)m.ref - (int64_t)m.qpos;
  a->pos = a->rcpos - (int64_t)ix.startpos[si];
  const uint32_t extra = (uint32_t)q - m.len - m.qpos;
  if (ix.rcref && (si & 1u)) {
    si -= 1;
    a->pos = (int64_t)ix.sizes[si] - a->pos - (int64_t)q;
    a->prefix = (uint16_t)extra; a->suffix = (uint16_t)m.qpos; a->rc = 1;
  } else {
    a->prefix = (uint16_t)m.qpos; a->suffix = (uint16_t)extra; a->rc = 0;
  }
  a->si = si; a->qpos = (uint16_t)m.qpos; a->len = (uint16_t)m.len;
  a->xm = 0; a->xu = 0; a->last = 0; a->first_item = 0; a->last_item = 0;
}

// prepare_matches (query.cpp:231-306) + set_nomap (308-320), lane-serial.
//   aln[n_in] scratch, ord[n_in] scratch; writes items[] (merge order), recs[] (HI order), sum.
// Returns the number of records (0 if the read prints nothing).
HDN inline int build_records(const DevIndex &ix, const Match *matches, int n_in, int q, int nomap,
                             Aln *aln, uint16_t *ord, Item *items, Rec *recs, ReadSum *sum) {
  int n = 0;
  for (int i = 0; i < n_in; ++i) {
    Aln a; resolve_match(ix, matches[i], q, &a);
    if (a.pos < 0) continue;                          // query.cpp:239-246
    aln[n] = a; ord[n] = (uint16_t)n; ++n;
  }
  sum->has_best = 0; sum->unmapped = 0; sum->best_pos = 0; sum->best_si = 0; sum->n_rec = 0;
  int n_rec = 0;
  if (n) {
    StdSort<LessMerge> s1{ord, LessMerge{aln}}; s1.run(n);
    int group_first = 0;
    for (int i = 0; i < n; ++i) {
      Aln &a = aln[ord[i]];
      items[i].prefix = a.prefix; items[i].len = a.len;
      a.xm += 1; a.xu += a.len;
      const bool end = (i + 1 == n) || aln[ord[i + 1]].pos != a.pos || aln[ord[i + 1]].si != a.si ||
                       aln[ord[i + 1]].rc != a.rc;
      if (end) {
        a.last = 1; a.first_item = (uint16_t)group_first; a.last_item = (uint16_t)i; group_first = i + 1;
      } else {
        Aln &na = aln[ord[i + 1]];
        na.qpos = a.qpos < na.qpos ? a.qpos : na.qpos;
        na.xm = a.xm; a.xm = 0;                       // ::swap + zero (query.cpp:284-287)
        na.xu = a.xu; a.xu = 0;
      }
    }
    StdSort<LessPrint> s2{ord, LessPrint{aln}}; s2.run(n);
    const Aln &b = aln[ord[0]];
    sum->has_best = 1; sum->best_si = b.si; sum->best_pos = b.pos;
    for (int i = 0; i < n; ++i) {
      const Aln &a = aln[ord[i]];
      if (!a.last) continue;
      Rec r;
      const int last_item = a.last_item;
      r.pos = a.pos; r.rcpos = a.rcpos; r.si = a.si; r.rc = a.rc;
      r.item_begin = a.first_item; r.item_cnt = (uint16_t)(last_item - a.first_item + 1);
      r.xu = a.xu; r.xe = 0;
      const int lead = items[a.first_item].prefix;
      const int endq = items[last_item].prefix + items[last_item].len;
      r.suffix = (uint16_t)(q - endq);
      r.qlen = (uint16_t)(endq - lead);
      r.L0 = 0; r.R0 = 0; r.flags = 0; r.seq_off = 0; r.lr_len = 0;
      recs[n_rec++] = r;
    }
  }
  if (n_rec == 0 && nomap) {                          // set_nomap
    sum->unmapped = 1; n_rec = 1;
    Rec r; r.pos = 0; r.rcpos = 0; r.si = 0; r.rc = 0; r.item_begin = 0; r.item_cnt = 0;
    r.xu = 0; r.xe = 0; r.qlen = 0; r.suffix = 0; r.L0 = 0; r.R0 = 0; r.flags = 0; r.seq_off = 0; r.lr_len = 0;
    recs[0] = r;
  }
  sum->n_rec = (uint16_t)n_rec;
  return n_rec;
}

// XE contribution of read bytes [j0, j0+8) of one record (query.cpp:270-274).
HD int xe_word(const DevIndex &ix, const uint8_t *P, int q, int64_t rcpos, int j0) {
  int cnt = 0;
  const int64_t rp = rcpos + j0;
  if (rp >= 0 && rp + 8 <= (int64_t)ix.N && j0 + 8 <= q) {
    uint64_t d = text8(ix.text, rp) ^ read8(P, j0);
    // count zero bytes of d
    uint64_t t = (d & 0x7f7f7f7f7f7f7f7fULL) + 0x7f7f7f7f7f7f7f7fULL;
    t = ~(t | d | 0x7f7f7f7f7f7f7f7fULL);
#if defined(__CUDA_ARCH__)
    cnt = __popcll(t);
#else
    cnt = __builtin_popcountll(t);
#endif
  } else {
    for (int j = j0; j < j0 + 8 && j < q; ++j) {
      const int64_t p = rcpos + j;
      if (p >= 0 && p < (int64_t)ix.N && ix.text[p] == P[j]) ++cnt;
    }
  }
  return cnt;
}

// mappability_tag.cpp:93-121 for one '=' block: abs = off32[chr] + POS (32-bit), block at read
// offset `off` of length `cnt`.  Returns false where the reference throws.
HD bool map_lr(const DevIndex &ix, uint32_t chrom, int64_t pos0, uint32_t off, uint32_t cnt,
               int *L, int *R) {
  const uint32_t abspos = ix.chrom_off32[chrom] + (uint32_t)(pos0 + 1);
  const uint32_t li = abspos + off + cnt - 1u, ri = abspos + off - 1u;
  const uint64_t lb = 2ull * li, rb = 2ull * ri + 1ull;
  const int lm = lb < ix.map_bytes ? ix.mapbody[lb] : 0;
  const int rm = rb < ix.map_bytes ? ix.mapbody[rb] : 0;
  *L = lm ? lm - 1 : 255;
  *R = rm ? rm : 255;
  return !((uint32_t)*L > cnt || (uint32_t)*R > cnt);
}

// ---- SAM text --------------------------------------------------------------------------------
// Every sink takes single chars, byte strings, and `word`: up to 8 chars packed little-endian in a
// register (first char in the low byte) -- the form literals and decimal numbers are produced in.
struct CountSink {
  uint32_t n = 0;
  HDN void ch(char) { ++n; }
  HDN void put(const char *, int len) { n += (uint32_t)len; }
  HDN void word(uint64_t, int len) { n += (uint32_t)len; }
};
struct BufSink {
  char *p; uint32_t n = 0;
  HDN void ch(char c) { p[n++] = c; }
  HDN void put(const char *s, int len) { for (int i = 0; i < len; ++i) p[n + i] = s[i]; n += (uint32_t)len; }
  HDN void word(uint64_t v, int len) { for (int i = 0; i < len; ++i) { p[n + i] = (char)v; v >>= 8; } n += (uint32_t)len; }
};
struct CapSink {
  char *p; uint32_t cap; uint32_t n = 0;
  HDN void ch(char c) { if (n < cap) p[n] = c; ++n; }
  HDN void put(const char *s, int len) { for (int i = 0; i < len; ++i) { if (n < cap) p[n] = s[i]; ++n; } }
  HDN void word(uint64_t v, int len) { for (int i = 0; i < len; ++i) { if (n < cap) p[n] = (char)v; ++n; v >>= 8; } }
};
// Byte stream -> HBM at ANY alignment with aligned 8-byte stores: characters are gathered in a register
// and flushed a word at a time; the partial words at both ends of the region (shared with the
// neighbouring column, which another kernel writes) are written byte by byte.
// the partial words at the two ends of a WordSink's region, byte by byte (kept out of line: the sink's `word` is inlined at
// every formatter call site and this loop would be copied into each)
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void wordsink_flush_partial(char *w, uint64_t acc, int lo, int fill) {
  for (int i = lo; i < fill; ++i) w[i] = (char)(acc >> (8 * i));
}
struct WordSink {
  char *w;            // aligned address of the word being filled
  uint64_t acc;
  int fill, lo;       // bytes gathered so far in this word; first byte of the word that belongs to us
  uint32_t n;
  HDN explicit WordSink(char *dst) : w(dst - ((uintptr_t)dst & 7)), acc(0), fill((int)((uintptr_t)dst & 7)), lo(fill), n(0) {}
  HDN void flush() {
    if (lo == 0 && fill == 8) *reinterpret_cast<uint64_t *>(w) = acc;
    else wordsink_flush_partial(w, acc, lo, fill);
    w += 8; acc = 0; fill = 0; lo = 0;
  }
  HDN void ch(char c) { acc |= (uint64_t)(uint8_t)c << (8 * fill); ++n; if (++fill == 8) flush(); }
  HDN void put(const char *s, int len) { for (int i = 0; i < len; ++i) ch(s[i]); }
  // up to 8 packed chars in O(1): merge into the word being filled, flush if it completes, keep the rest
  HDN void word(uint64_t v, int len) {
    if (len <= 0) return;
    n += (uint32_t)len;
    acc |= v << (8 * fill);
    const int room = 8 - fill;
    if (len >= room) {
      fill = 8; flush();
      if (len > room) { acc = v >> (8 * room); fill = len - room; }
    } else fill += len;
  }
  HDN void finish() { if (fill > lo) flush(); }
};
// number of decimal digits of v, without a digit loop: all the counting pass (k_sizes, CountSink) needs of a number
HDN inline int ndigits32(uint32_t v) {
  if (v < 100u) return v < 10u ? 1 : 2;
  if (v < 10000u) return v < 1000u ? 3 : 4;
  return 5 + (int)(v >= 100000u) + (int)(v >= 1000000u) + (int)(v >= 10000000u) + (int)(v >= 100000000u) + (int)(v >= 1000000000u);
}
HDN inline int ndigits64(uint64_t v) {
  if (v <= 0xffffffffull) return ndigits32((uint32_t)v);
  int n = 10; v /= 10000000000ull;
  while (v) { ++n; v /= 10u; }
  return n;
}
// up to 8 decimal digits of x as packed chars (most significant digit in the low byte).  The writing sinks keep the
// digit loop: a loop-free variant (two 4-digit halves split by multiply-shift into byte lanes, with a short path for
// one and two digits) was measured SLOWER inside k_emit_text (1.35 -> 1.53 ms inlined, 1.40 out of line, r02o A/B):
// its code is larger at every call site of the field loop.
HDN inline uint64_t digits8(uint32_t x, bool pad8, int *nd) {
  uint64_t packed = 0; int n = 0;
  do { const uint32_t d = x / 10u; packed = (packed << 8) | (uint64_t)('0' + (x - d * 10u)); x = d; ++n; } while (x || (pad8 && n < 8));
  *nd = n;
  return packed;
}
template <class S> HDN inline void put_u64(S &s, uint64_t v) {
  int nd;
  if (v < 100000000ull) {                        // the common case: at most 8 digits, one packed word
    const uint64_t p = digits8((uint32_t)v, false, &nd);
    s.word(p, nd);
    return;
  }
  // longer numbers: leading chunk unpadded, the following 8-digit chunks zero padded (no recursion)
  if (v <= 0xffffffffull) {                      // 9 or 10 digits (positions beyond 10^8): 32-bit arithmetic
    const uint32_t top = (uint32_t)v / 100000000u, low = (uint32_t)v - top * 100000000u;
    { const uint64_t p = digits8(top, false, &nd); s.word(p, nd); }
    { const uint64_t p = digits8(low, true, &nd); s.word(p, nd); }
    return;
  }
  const uint64_t top = v / 10000000000000000ull, rest = v % 10000000000000000ull;
  const uint32_t mid = (uint32_t)(rest / 100000000ull), low = (uint32_t)(rest % 100000000ull);
  if (top) { const uint64_t p = digits8((uint32_t)top, false, &nd); s.word(p, nd); }
  { const uint64_t p = digits8(mid, top != 0, &nd); s.word(p, nd); }
  { const uint64_t p = digits8(low, true, &nd); s.word(p, nd); }
}
// the counting sink needs the number of digits only
HDN inline void put_u64(CountSink &s, uint64_t v) { s.n += (uint32_t)ndigits64(v); }
template <class S> HDN inline void put_i64(S &s, int64_t v) {
  if (v < 0) { s.ch('-'); put_u64(s, (uint64_t)(-v)); } else put_u64(s, (uint64_t)v);
}
// string literals are packed at compile time, 8 chars per word
HDN constexpr uint64_t pack8(const char *s, int n) {
  uint64_t v = 0;
  for (int i = n - 1; i >= 0; --i) v = (v << 8) | (uint64_t)(uint8_t)s[i];
  return v;
}
template <class S, int N> HDN inline void put_lit(S &s, const char (&lit)[N]) {
  constexpr int len = N - 1;
  if (len <= 8) { s.word(pack8(lit, len), len); }
  else { s.word(pack8(lit, 8), 8); s.word(pack8(lit + 8, len - 8 > 8 ? 8 : len - 8), len - 8 > 8 ? 8 : len - 8); }
  static_assert(N - 1 <= 16, "literal too long for put_lit");
}
template <class S> HDN inline void put_descr(S &s, const DevIndex &ix, uint32_t si) {
  const int len = ix.descr_off[si + 1] - ix.descr_off[si];
  if (len <= 8 && ix.descr8) s.word(ix.descr8[si], len);           // names of <= 8 chars: one packed word
  else s.put(ix.descr + ix.descr_off[si], len);
}
// CIGAR of a record (query.cpp:260-268): [<prefix>S] len= [<gap>M len=]... [<suffix>S].  One loop over its
// <number><letter> tokens, so that the number formatter exists ONCE in the emitted code (see put_fields).
template <class S> HDN inline void put_cigar(S &s, const Rec &r, const Item *items) {
  if (r.item_cnt == 0) { s.ch('*'); return; }
  uint32_t last_end = 0;
  const int n_tok = 2 * (int)r.item_cnt + 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int t = 0; t < n_tok; ++t) {
    uint32_t num; char c; bool present;
    if (t == n_tok - 1) { num = r.suffix; c = 'S'; present = r.suffix != 0; }
    else {
      const Item it = items[r.item_begin + (t >> 1)];
      if (t & 1) { num = it.len; c = '='; present = true; last_end = (uint32_t)it.prefix + it.len; }
      else { num = it.prefix - last_end; c = last_end ? 'M' : 'S'; present = it.prefix != 0; }
    }
    if (!present) continue;
    put_u64(s, num); s.ch(c);
  }
}

struct MateView { uint8_t has; uint32_t si; int64_t pos; };

// set_mate (query.cpp:421-434) seen from one read: which (chr,pos) goes into RNEXT/PNEXT (and,
// for a placeholder, RNAME/POS), and whether flag 8 is added.
HD void mate_view(uint16_t my_flag, const ReadSum &me, uint16_t other_flag, const ReadSum *other,
                  bool is_first_of_pair, MateView *mv, uint16_t *flag_out) {
  mv->has = 0; mv->si = 0; mv->pos = 0; *flag_out = my_flag;
  if (!other) return;
  const uint16_t f1 = is_first_of_pair ? my_flag : other_flag;
  const uint16_t f2 = is_first_of_pair ? other_flag : my_flag;
  if (!((f1 & 64) && (f2 & 128))) return;              // has_mate (query.cpp:417-419)
  if (!(me.n_rec && other->n_rec)) return;
  if (other->has_best) { mv->has = 1; mv->si = other->best_si; mv->pos = other->best_pos; }
  else {
    *flag_out = (uint16_t)(my_flag | 8);
    if (me.has_best) { mv->has = 1; mv->si = me.best_si; mv->pos = me.best_pos; }
  }
}

// A SAM line's computed text is a sequence of FIELDS: a literal of up to 8 characters followed by one value (number,
// sequence name, character, CIGAR).  put_head / put_tags walk the field list in ONE loop whose body holds the only copy
// of every formatter: the kernels stay small enough for the instruction cache (the fully inlined version was 35 k
// instructions and spent 30 % of its issue slots waiting for instruction fetches), and the threads of a warp are at the
// same field at the same time whatever their records look like.
enum { FK_NONE = 0, FK_U64, FK_I64, FK_DESCR, FK_CHAR, FK_CIGAR };
struct Field { uint64_t lit; int lit_len; int kind; int64_t val; const Rec *rec; };
constexpr int F_HEAD_END = 7, F_TAGS_END = 21;
#define SMASH_LIT(str) f.lit = pack8(str, (int)sizeof(str) - 1); f.lit_len = (int)sizeof(str) - 1
HD Field line_field(int fi, uint16_t flag, bool unmapped, const Rec *recs, int hi, int n_rec, const MateView &mv) {
  Field f{0, 0, FK_NONE, 0, nullptr};
  const Rec &r = recs[hi];
  if (fi < 4) {                                                // FLAG RNAME POS MAPQ CIGAR (query.cpp:331-372)
    if (unmapped) {
      switch (fi) {
        case 0: SMASH_LIT("\t"); f.kind = FK_U64; f.val = flag; break;
        case 1: if (mv.has) { SMASH_LIT("\t"); f.kind = FK_DESCR; f.val = mv.si; } else { SMASH_LIT("\t*\t0"); } break;
        case 2: if (mv.has) { SMASH_LIT("\t"); f.kind = FK_I64; f.val = mv.pos + 1; } break;
        default: SMASH_LIT("\t0\t*"); break;
      }
    } else {
      switch (fi) {
        case 0: SMASH_LIT("\t"); f.kind = FK_U64; f.val = (int64_t)(flag | (r.rc ? 16 : 0) | (hi ? 256 : 0)); break;
        case 1: SMASH_LIT("\t"); f.kind = FK_DESCR; f.val = r.si; break;
        case 2: SMASH_LIT("\t"); f.kind = FK_I64; f.val = r.pos + 1; break;
        default: SMASH_LIT("\t50\t"); f.kind = FK_CIGAR; f.rec = &r; break;
      }
    }
  } else if (fi < F_HEAD_END) {                                // RNEXT PNEXT TLEN
    if (mv.has) {
      switch (fi) {
        case 4: SMASH_LIT("\t"); f.kind = FK_DESCR; f.val = mv.si; break;
        case 5: SMASH_LIT("\t"); f.kind = FK_I64; f.val = mv.pos + 1; break;
        default: SMASH_LIT("\t0\t"); break;
      }
    } else if (fi == 4) { SMASH_LIT("\t*\t0\t0\t"); }
  } else if (unmapped) {                                       // tags of a placeholder
    if (fi == 7) { SMASH_LIT("\tXM:i:0"); } else if (fi == 8) { SMASH_LIT("\tNH:i:0"); }
  } else if (fi < 13) {
    switch (fi) {
      case 7: SMASH_LIT("\tXM:i:"); f.kind = FK_U64; f.val = r.item_cnt; break;
      case 8: SMASH_LIT("\tXU:i:"); f.kind = FK_U64; f.val = r.xu; break;
      case 9: SMASH_LIT("\tXE:i:"); f.kind = FK_U64; f.val = r.xe; break;
      case 10: SMASH_LIT("\tXS:A:"); f.kind = FK_CHAR; f.val = r.rc ? '-' : '+'; break;
      case 11: SMASH_LIT("\tNH:i:"); f.kind = FK_U64; f.val = n_rec; break;
      default: SMASH_LIT("\tHI:i:"); f.kind = FK_U64; f.val = hi; break;
    }
  } else if (fi < 17) {                                        // previous record of the read (lower-case tags)
    if (hi > 0) {
      const Rec &p = recs[hi - 1];
      switch (fi) {
        case 13: SMASH_LIT("\tcc:Z:"); f.kind = FK_DESCR; f.val = p.si; break;
        case 14: SMASH_LIT("\tcp:i:"); f.kind = FK_I64; f.val = p.pos + 1; break;
        case 15: SMASH_LIT("\txo:A:"); f.kind = FK_CHAR; f.val = p.rc == r.rc ? '=' : '!'; break;
        default: SMASH_LIT("\txc:Z:"); f.kind = FK_CIGAR; f.rec = &p; break;
      }
    }
  } else if (hi + 1 < n_rec) {                                 // next record (upper-case tags)
    const Rec &x = recs[hi + 1];
    switch (fi) {
      case 17: SMASH_LIT("\tCC:Z:"); f.kind = FK_DESCR; f.val = x.si; break;
      case 18: SMASH_LIT("\tCP:i:"); f.kind = FK_I64; f.val = x.pos + 1; break;
      case 19: SMASH_LIT("\tXO:A:"); f.kind = FK_CHAR; f.val = x.rc == r.rc ? '=' : '!'; break;
      default: SMASH_LIT("\tXC:Z:"); f.kind = FK_CIGAR; f.rec = &x; break;
    }
  }
  return f;
}
#undef SMASH_LIT
template <class S>
HDN inline void put_fields(S &s, const DevIndex &ix, int f0, int f1, uint16_t flag, bool unmapped, const Rec *recs, int hi,
                           int n_rec, const Item *items, const MateView &mv) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int fi = f0; fi < f1; ++fi) {
    const Field f = line_field(fi, flag, unmapped, recs, hi, n_rec, mv);
    if (!f.lit_len) continue;
    s.word(f.lit, f.lit_len);
    if (f.kind == FK_U64 || f.kind == FK_I64) {
      uint64_t v = (uint64_t)f.val;
      if (f.kind == FK_I64 && f.val < 0) { s.ch('-'); v = (uint64_t)(-f.val); }
      put_u64(s, v);
    } else if (f.kind == FK_DESCR) put_descr(s, ix, (uint32_t)f.val);
    else if (f.kind == FK_CHAR) s.ch((char)f.val);
    else if (f.kind == FK_CIGAR) put_cigar(s, *f.rec, items);
  }
}
// Columns 1-9 (up to and including the tab before SEQ).  `recs` = the read's records, `hi` = this one.
template <class S>
HDN inline void put_head(S &s, const DevIndex &ix, const char *name, int name_len, uint16_t flag,
                         bool unmapped, const Rec *recs, int hi, int n_rec, const Item *items, const MateView &mv) {
  s.put(name, name_len);
  put_fields(s, ix, 0, F_HEAD_END, flag, unmapped, recs, hi, n_rec, items, mv);
}
// Everything after QUAL up to (not including) the optional fields / L,R tags / newline.
template <class S>
HDN inline void put_tags(S &s, const DevIndex &ix, bool unmapped, const Rec *recs, int hi, int n_rec,
                         const Item *items) {
  const MateView none{0, 0, 0};
  put_fields(s, ix, F_HEAD_END, F_TAGS_END, 0, unmapped, recs, hi, n_rec, items, none);
}
// mappability_tag's appended tags (after the optional fields): \tL<u>:i:x\tR<u>:i:y for u < 10.
template <class S>
HDN inline bool put_lr_tags(S &s, const DevIndex &ix, const Rec &r, const Item *items) {
  const int n = r.item_cnt < 10 ? (int)r.item_cnt : 10;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int t = 0; t < 2 * n; ++t) {
    const Item it = items[r.item_begin + (t >> 1)];            // L/R were looked up once, by k_rec_xe (map_lr)
    // "\tL<u>:i:" as one packed word
    const uint64_t lit = (uint64_t)'\t' | ((uint64_t)((t & 1) ? 'R' : 'L') << 8) | ((uint64_t)('0' + (t >> 1)) << 16) |
                         ((uint64_t)':' << 24) | ((uint64_t)'i' << 32) | ((uint64_t)':' << 40);
    s.word(lit, 6);
    put_u64(s, (uint64_t)((t & 1) ? it.R : it.L));
  }
  return true;
}

// reverse_complement (fasta.cpp:26-61) for one character
HD uint8_t comp_char(uint8_t c) {
  switch (c) {
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
    case 'r': return 'y'; case 'y': return 'r'; case 'm': return 'k'; case 'k': return 'm';
    case 'b': return 'v'; case 'd': return 'h'; case 'h': return 'd'; case 'v': return 'b';
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'M': return 'K'; case 'K': return 'M';
    case 'B': return 'V'; case 'D': return 'H'; case 'H': return 'D'; case 'V': return 'B';
    default: return c;
  }
}

}  // namespace smash
