/* oracle/smash_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see smash_oracle.h).
 *
 * CPU restatement, in plain C, of the reference's hot path.  Every function names the reference
 * lines it restates (paths relative to /root/reference).  The control flow of the two search
 * modes is kept faithful on purpose: MEM mode's output depends on it.
 */
#define _GNU_SOURCE
#include "smash_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ index access */

static inline uint64_t sa_at(const orc_index *ix, uint64_t i) {
  return ix->w == 4 ? ((const uint32_t *)ix->sa)[i] : ((const uint64_t *)ix->sa)[i];
}
static inline uint64_t isa_at(const orc_index *ix, uint64_t i) {
  return ix->w == 4 ? ((const uint32_t *)ix->isa)[i] : ((const uint64_t *)ix->isa)[i];
}
/* vec_uchar::operator[] (longSA.h:34-39): byte value, 255 => lower_bound in the sorted table. */
static inline uint64_t lcp_at(const orc_index *ix, uint64_t i) {
  uint8_t v = ix->lcp_vec[i];
  if (v != 255) return v;
  uint64_t lo = 0, hi = ix->n_m;
  while (lo < hi) {
    uint64_t mid = lo + (hi - lo) / 2;
    if (ix->lcp_m[mid].idx < i) lo = mid + 1; else hi = mid;
  }
  uint64_t val = ix->lcp_m[lo].val;
  return ix->w == 4 ? (val & 0xffffffffu) : val;
}
static inline uint64_t log_n(const orc_index *ix) {        /* longSA.cpp:97 */
  return (uint64_t)ceil(log((double)ix->N) / log(2.0));
}

typedef struct { uint64_t depth, lo, hi; } ivl;            /* interval_t, longSA.h:64-75 */

/* top_down_faster (longSA.cpp:322-380): narrow [lo,hi] (all suffixes share `off` chars) to the
 * suffixes whose next char is c.  The reference's hand-rolled double binary search computes
 * exactly the first/last such index; we state that directly.  Returns 0 if none. */
static int narrow(const orc_index *ix, uint8_t c, uint64_t off, uint64_t *lo, uint64_t *hi) {
  const uint8_t *t = ix->text;
  uint64_t a = *lo, b = *hi;
  if (c < t[sa_at(ix, a) + off] || c > t[sa_at(ix, b) + off]) return 0;
  uint64_t l = a, r = b + 1;                   /* first index with char >= c */
  while (l < r) {
    uint64_t m = l + (r - l) / 2;
    if (t[sa_at(ix, m) + off] < c) l = m + 1; else r = m;
  }
  uint64_t first = l;
  if (first > b || t[sa_at(ix, first) + off] != c) return 0;
  l = first; r = b + 1;                        /* first index with char > c */
  while (l < r) {
    uint64_t m = l + (r - l) / 2;
    if (t[sa_at(ix, m) + off] <= c) l = m + 1; else r = m;
  }
  *lo = first; *hi = l - 1;
  return 1;
}

/* traverse (longSA.cpp:297-316) */
static void traverse(const orc_index *ix, const uint8_t *P, uint64_t q, uint64_t prefix, ivl *cur,
                     uint64_t stop_len) {
  if (cur->depth >= stop_len) return;
  while (prefix + cur->depth < q) {
    uint64_t lo = cur->lo, hi = cur->hi;
    if (!narrow(ix, P[prefix + cur->depth], cur->depth, &lo, &hi)) return;
    cur->depth += 1; cur->lo = lo; cur->hi = hi;
    if (cur->depth == stop_len) return;
  }
}

/* expand_link (longSA.h:158-174): grow an ISA-derived interval by LCP, giving up after
 * 2*depth*logN steps. */
static int expand_link(const orc_index *ix, ivl *v, uint64_t logN) {
  const uint64_t thresh = 2 * v->depth * logN;
  uint64_t steps = 0, lo = v->lo, hi = v->hi;
  while (lcp_at(ix, lo) >= v->depth) { if (++steps >= thresh) return 0; --lo; }
  while (hi < ix->N - 1 && lcp_at(ix, hi + 1) >= v->depth) { if (++steps >= thresh) return 0; ++hi; }
  v->lo = lo; v->hi = hi;
  return 1;
}

/* suffixlink (longSA.cpp:383-392) */
static int suffixlink(const orc_index *ix, ivl *v, uint64_t logN) {
  if (v->depth <= 1) { v->depth = 0; return 0; }
  v->depth -= 1;
  v->lo = isa_at(ix, sa_at(ix, v->lo) + 1);
  v->hi = isa_at(ix, sa_at(ix, v->hi) + 1);
  return expand_link(ix, v, logN);
}

typedef struct { orc_match *out; uint64_t cap, n; } sink;
static inline void emit(sink *s, uint64_t ref, uint64_t query, uint64_t len) {
  if (s->n < s->cap) { s->out[s->n].ref = ref; s->out[s->n].query = query; s->out[s->n].len = len; }
  s->n++;
}

/* longSA::MAM (longSA.cpp:503-536) with is_leftmaximal (longSA.cpp:540-546) */
static void mam_search(const orc_index *ix, const uint8_t *P, uint64_t q, uint64_t min_len, sink *s) {
  const uint64_t logN = log_n(ix);
  ivl cur = {0, 0, ix->N - 1};
  uint64_t prefix = 0;
  while (prefix < q) {
    traverse(ix, P, q, prefix, &cur, q);
    if (cur.depth <= 1) { cur.depth = 0; cur.lo = 0; cur.hi = ix->N - 1; ++prefix; continue; }
    if (cur.hi == cur.lo && cur.depth >= min_len) {
      uint64_t r = sa_at(ix, cur.lo);
      if (prefix == 0 || r == 0 || P[prefix - 1] != ix->text[r - 1]) emit(s, r, prefix, cur.depth);
    }
    do {
      cur.depth -= 1;
      cur.lo = isa_at(ix, sa_at(ix, cur.lo) + 1);
      cur.hi = isa_at(ix, sa_at(ix, cur.hi) + 1);
      ++prefix;
      if (cur.depth == 0 || !expand_link(ix, &cur, logN)) {
        cur.depth = 0; cur.lo = 0; cur.hi = ix->N - 1;
        break;
      }
    } while (cur.depth > 0 && cur.hi == cur.lo);
  }
}

/* find_Lmaximal (longSA.cpp:438-457) */
static inline void left_maximal_emit(const orc_index *ix, const uint8_t *P, uint64_t min_len,
                                     uint64_t prefix, uint64_t r, uint64_t len, sink *s) {
  if (prefix == 0 || r == 0 || P[prefix - 1] != ix->text[r - 1])
    if (len >= min_len) emit(s, r, prefix, len);
}

/* collectMEMs (longSA.cpp:461-490): both intervals by value. */
static void collect_mems(const orc_index *ix, const uint8_t *P, uint64_t min_len, uint64_t prefix,
                         ivl mli, ivl xmi, sink *s) {
  for (uint64_t i = xmi.lo; i <= xmi.hi; ++i)
    left_maximal_emit(ix, P, min_len, prefix, sa_at(ix, i), xmi.depth, s);
  if (mli.lo == xmi.lo && mli.hi == xmi.hi) return;
  while (xmi.depth >= mli.depth) {
    if (xmi.hi + 1 < ix->N) {
      uint64_t a = lcp_at(ix, xmi.lo), b = lcp_at(ix, xmi.hi + 1);
      xmi.depth = a > b ? a : b;
    } else {
      xmi.depth = lcp_at(ix, xmi.lo);
    }
    if (xmi.depth >= mli.depth) {
      while (lcp_at(ix, xmi.lo) >= xmi.depth) {
        --xmi.lo;
        left_maximal_emit(ix, P, min_len, prefix, sa_at(ix, xmi.lo), xmi.depth, s);
      }
      while (xmi.hi + 1 < ix->N && lcp_at(ix, xmi.hi + 1) >= xmi.depth) {
        ++xmi.hi;
        left_maximal_emit(ix, P, min_len, prefix, sa_at(ix, xmi.hi), xmi.depth, s);
      }
    }
  }
}

/* findMEM (longSA.cpp:395-435) -- note prefix starts at 1 and the ignored suffixlink(&xmi). */
static void mem_search(const orc_index *ix, const uint8_t *P, uint64_t q, uint64_t min_len, sink *s) {
  if (min_len < 1) return;                                   /* longSA::MEM, longSA.cpp:587-590 */
  const uint64_t logN = log_n(ix);
  const uint64_t last = ix->N - 1;
  uint64_t prefix = 1;
  ivl mli = {0, 0, last}, xmi = {0, 0, last};
  while (prefix <= q) {
    traverse(ix, P, q, prefix, &mli, min_len);
    if (mli.depth > xmi.depth) xmi = mli;
    if (mli.depth <= 1) {
      mli.depth = 0; mli.lo = 0; mli.hi = last; xmi = mli;
      ++prefix;
      continue;
    }
    if (mli.depth >= min_len) {
      traverse(ix, P, q, prefix, &xmi, q);
      collect_mems(ix, P, min_len, prefix, mli, xmi, s);
      ++prefix;
      if (!suffixlink(ix, &mli, logN)) {
        mli.depth = 0; mli.lo = 0; mli.hi = last; xmi = mli;
        continue;
      }
      (void)suffixlink(ix, &xmi, logN);
    } else {
      ++prefix;
      if (!suffixlink(ix, &mli, logN)) {
        mli.depth = 0; mli.lo = 0; mli.hi = last; xmi = mli;
        continue;
      }
      xmi = mli;
    }
  }
}

uint64_t orc_mam(const orc_index *ix, const uint8_t *query, uint64_t qlen, uint64_t min_len,
                 orc_match *out, uint64_t cap) {
  sink s = {out, cap, 0};
  mam_search(ix, query, qlen, min_len, &s);
  return s.n;
}
uint64_t orc_mem(const orc_index *ix, const uint8_t *query, uint64_t qlen, uint64_t min_len,
                 orc_match *out, uint64_t cap) {
  sink s = {out, cap, 0};
  mem_search(ix, query, qlen, min_len, &s);
  return s.n;
}

/* SURVEY.md Appendix A.1, stated without any index: for every p, the longest prefix of P[p..]
 * occurring in T, its occurrence count, and the left-maximality test. */
uint64_t orc_mam_bruteforce(const uint8_t *T, uint64_t N, const uint8_t *P, uint64_t q,
                            uint64_t min_len, orc_match *out, uint64_t cap) {
  sink s = {out, cap, 0};
  for (uint64_t p = 0; p < q; ++p) {
    uint64_t best = 0, cnt = 0, where = 0;
    for (uint64_t r = 0; r < N; ++r) {
      uint64_t l = 0;
      while (p + l < q && r + l < N && T[r + l] == P[p + l]) ++l;
      if (l > best) { best = l; cnt = 1; where = r; }
      else if (l == best && l > 0) ++cnt;
    }
    if (best >= 2 && cnt == 1 && best >= min_len &&
        (p == 0 || where == 0 || P[p - 1] != T[where - 1]))
      emit(&s, where, p, best);
  }
  return s.n;
}

/* ------------------------------------------------------------------ libstdc++ std::sort
 * query.cpp:251,290 call std::sort on vector<Alignment*>; in MEM mode to_print has real ties,
 * so HI numbering depends on libstdc++'s introsort (bits/stl_algo.h: threshold 16, median of
 * three to first, unguarded partition, final insertion sort).  Restated for an index array. */
typedef int (*less_fn)(const void *ctx, int a, int b);
typedef struct { int *v; less_fn lt; const void *ctx; } sorter;
#define LT(S, a, b) ((S)->lt((S)->ctx, (a), (b)))

static void s_unguarded_linear_insert(sorter *S, int last) {
  int *v = S->v, val = v[last], next = last - 1;
  while (LT(S, val, v[next])) { v[last] = v[next]; last = next; --next; }
  v[last] = val;
}
static void s_insertion_sort(sorter *S, int first, int last) {
  int *v = S->v;
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (LT(S, v[i], v[first])) {
      int val = v[i];
      memmove(&v[first + 1], &v[first], (size_t)(i - first) * sizeof(int));
      v[first] = val;
    } else {
      s_unguarded_linear_insert(S, i);
    }
  }
}
static void s_push_heap(sorter *S, int first, int hole, int top, int val) {
  int *v = S->v, parent = (hole - 1) / 2;
  while (hole > top && LT(S, v[first + parent], val)) {
    v[first + hole] = v[first + parent];
    hole = parent; parent = (hole - 1) / 2;
  }
  v[first + hole] = val;
}
static void s_adjust_heap(sorter *S, int first, int hole, int len, int val) {
  int *v = S->v;
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (LT(S, v[first + child], v[first + child - 1])) child--;
    v[first + hole] = v[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[first + hole] = v[first + child - 1];
    hole = child - 1;
  }
  s_push_heap(S, first, hole, top, val);
}
static void s_heap_sort(sorter *S, int first, int last) {       /* __partial_sort(first,last,last) */
  int *v = S->v, len = last - first;
  if (len >= 2)
    for (int parent = (len - 2) / 2;; --parent) {
      s_adjust_heap(S, first, parent, len, v[first + parent]);
      if (parent == 0) break;
    }
  while (last - first > 1) {
    --last;
    int val = v[last];
    v[last] = v[first];
    s_adjust_heap(S, first, 0, last - first, val);
  }
}
static void s_median_to_first(sorter *S, int res, int a, int b, int c) {
  int *v = S->v, t;
#define SWP(x, y) (t = v[x], v[x] = v[y], v[y] = t)
  if (LT(S, v[a], v[b])) {
    if (LT(S, v[b], v[c])) SWP(res, b); else if (LT(S, v[a], v[c])) SWP(res, c); else SWP(res, a);
  } else if (LT(S, v[a], v[c])) SWP(res, a);
  else if (LT(S, v[b], v[c])) SWP(res, c);
  else SWP(res, b);
}
static int s_unguarded_partition(sorter *S, int first, int last, int pivot) {
  int *v = S->v, t;
  for (;;) {
    while (LT(S, v[first], v[pivot])) ++first;
    --last;
    while (LT(S, v[pivot], v[last])) --last;
    if (!(first < last)) return first;
    SWP(first, last);
    ++first;
  }
}
static void s_introsort_loop(sorter *S, int first, int last, int depth) {
  while (last - first > 16) {
    if (depth == 0) { s_heap_sort(S, first, last); return; }
    --depth;
    int mid = first + (last - first) / 2;
    s_median_to_first(S, first, first + 1, mid, last - 1);
    int cut = s_unguarded_partition(S, first + 1, last, first);
    s_introsort_loop(S, cut, last, depth);
    last = cut;
  }
}
static void std_sort(int *v, int n, less_fn lt, const void *ctx) {
  if (n <= 0) return;
  sorter S = {v, lt, ctx};
  int lg = 0;
  for (int k = n; k > 1; k >>= 1) ++lg;
  s_introsort_loop(&S, 0, n, 2 * lg);
  if (n > 16) {
    s_insertion_sort(&S, 0, 16);
    for (int i = 16; i != n; ++i) s_unguarded_linear_insert(&S, i);
  } else {
    s_insertion_sort(&S, 0, n);
  }
}

/* longSA::MUM (longSA.cpp:549-585): MAM matches, sorted by (ref asc, len desc) with std::sort, then the
 * MUMmer-3 cleanMUMcand sweep drops matches contained in / ending with an earlier one.  Survivors are
 * handed on in that sorted order.  m[0..n) in place; returns the new count. */
static int lt_by_ref(const void *ctx, int x, int y) {
  const orc_match *M = (const orc_match *)ctx;
  if (M[x].ref == M[y].ref) return M[x].len > M[y].len;
  return M[x].ref < M[y].ref;
}
static uint64_t mum_clean(orc_match *m, uint64_t n) {
  if (!n) return 0;
  int *ord = (int *)malloc(sizeof(int) * n);
  orc_match *tmp = (orc_match *)malloc(sizeof(orc_match) * n);
  for (uint64_t i = 0; i < n; ++i) ord[i] = (int)i;
  std_sort(ord, (int)n, lt_by_ref, m);
  for (uint64_t i = 0; i < n; ++i) tmp[i] = m[ord[i]];
  uint64_t out = 0, dbright = 0;
  int ignoreprevious = 0;
  for (uint64_t i = 0; i < n; ++i) {
    int ignorecurrent = 0;
    const uint64_t currentright = tmp[i].ref + tmp[i].len - 1;
    if (dbright > currentright) ignorecurrent = 1;
    else if (dbright == currentright) {
      ignorecurrent = 1;
      if (!ignoreprevious && i > 0 && tmp[i - 1].ref == tmp[i].ref) ignoreprevious = 1;
    } else dbright = currentright;
    if (i > 0 && !ignoreprevious) m[out++] = tmp[i - 1];
    ignoreprevious = ignorecurrent;
  }
  if (!ignoreprevious) m[out++] = tmp[n - 1];
  free(ord); free(tmp);
  return out;
}

/* ------------------------------------------------------------------ records (query.cpp) */

typedef struct {
  int64_t rcpos, pos, qpos;          /* Alignment, query.h:18-45 */
  uint64_t seq_index, prefix, length, suffix;
  uint64_t n_matches, n_unique, n_matched, hi_index;
  int prev, next, rc;
  int cigar_buf;                     /* which cigar buffer this alignment currently owns */
} aln;

typedef struct {
  aln *a; int n, cap_a;
  int *order;
  char *cig; int cig_stride;         /* cig[k*stride..] buffers that get swapped between alns */
  orc_match *m; uint64_t m_cap;
  uint8_t *lower; uint64_t lower_cap;
  int n_aln;                         /* n_alignments */
  int best;                          /* index into a[] or -1 */
  unsigned read_flag;
  int unmapped_placeholder;
  uint64_t n_found;                  /* matches as emitted by the search (before pos<0 erase) */
} read_state;

static int lt_merge(const void *ctx, int x, int y) {          /* to_merge, query.cpp:203-219 */
  const aln *A = (const aln *)ctx; const aln *a = &A[x], *b = &A[y];
  if (a->rc != b->rc) return a->rc < b->rc;
  if (a->seq_index != b->seq_index) return a->seq_index < b->seq_index;
  if (a->pos != b->pos) return a->pos < b->pos;
  return a->prefix < b->prefix;
}
static int lt_print(const void *ctx, int x, int y) {          /* to_print, query.cpp:221-229 */
  const aln *A = (const aln *)ctx; const aln *a = &A[x], *b = &A[y];
  if (a->qpos == b->qpos) return a->rc < b->rc;
  return a->qpos < b->qpos;
}

static void rs_reserve(read_state *r, int n, uint64_t q) {
  if (n + 1 > r->cap_a || (int)(q * 5 + 32) > r->cig_stride) {
    r->cap_a = n + 8;
    r->cig_stride = (int)(q * 5 + 32);
    r->a = (aln *)realloc(r->a, sizeof(aln) * (size_t)r->cap_a);
    r->order = (int *)realloc(r->order, sizeof(int) * (size_t)r->cap_a);
    r->cig = (char *)realloc(r->cig, (size_t)r->cap_a * (size_t)r->cig_stride);
  }
}

/* Aligner::run (query.cpp:322-329) = search + prepare_matches (231-306) + set_nomap (308-320) */
static void run_read(const orc_index *ix, const orc_params *p, const uint8_t *seq, uint64_t q,
                     unsigned read_flag, read_state *r) {
  if (q + 1 > r->lower_cap) { r->lower_cap = q + 64; r->lower = (uint8_t *)realloc(r->lower, r->lower_cap); }
  for (uint64_t i = 0; i < q; ++i) {                          /* NewQuery::extend, query.cpp:125-144 */
    uint8_t c = seq[i];
    if (c >= 'A' && c <= 'Z') c = (uint8_t)(c + 32);
    if (p->nucleotides_only && c != 'a' && c != 'c' && c != 'g' && c != 't') c = '~';
    r->lower[i] = c;
  }
  sink s = {r->m, r->m_cap, 0};
  for (;;) {
    s.n = 0;
    if (p->mode == ORC_MEM) mem_search(ix, r->lower, q, p->min_len, &s);
    else mam_search(ix, r->lower, q, p->min_len, &s);
    if (s.n <= s.cap) break;
    r->m_cap = s.n + 16; r->m = (orc_match *)realloc(r->m, sizeof(orc_match) * r->m_cap);
    s.out = r->m; s.cap = r->m_cap;
  }
  if (p->mode == ORC_MUM) s.n = mum_clean(r->m, s.n);
  r->n_found = s.n;
  r->read_flag = read_flag; r->n = 0; r->n_aln = 0; r->best = -1; r->unmapped_placeholder = 0;
  rs_reserve(r, (int)s.n, q);
  /* Alignment::resolve (query.cpp:68-97) + erase of pos<0 (query.cpp:239-246) */
  int kept = 0;
  for (uint64_t k = 0; k < s.n; ++k) {
    const orc_match *mt = &r->m[k];
    uint64_t lo = 0, hi = ix->n_descr;                        /* upper_bound(startpos, ref) */
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (ix->startpos[mid] <= mt->ref) lo = mid + 1; else hi = mid; }
    uint64_t si = lo - 1;
    aln a; memset(&a, 0, sizeof a);
    a.rcpos = (int64_t)mt->ref - (int64_t)mt->query;
    a.pos = a.rcpos - (int64_t)ix->startpos[si];
    unsigned extra = (unsigned)(q - mt->len - mt->query);
    if (ix->rcref && (si % 2) == 1) {
      si -= 1;
      a.pos = (int64_t)ix->sizes[si] - a.pos - (int64_t)q;
      a.prefix = extra; a.suffix = mt->query; a.rc = 1;
    } else {
      a.prefix = mt->query; a.suffix = extra; a.rc = 0;
    }
    a.seq_index = si; a.qpos = (int64_t)mt->query; a.length = mt->len;
    a.prev = a.next = -1;
    if (a.pos < 0) continue;
    a.cigar_buf = kept;
    r->a[kept] = a; r->order[kept] = kept;
    r->cig[(size_t)kept * r->cig_stride] = '*'; r->cig[(size_t)kept * r->cig_stride + 1] = 0;
    ++kept;
  }
  r->n = kept;
  if (kept) {                                                 /* sam_out is always true here */
    std_sort(r->order, kept, lt_merge, r->a);
    uint64_t cig_end = 0, last_end = 0;
    for (int i = 0; i < kept; ++i) {
      aln *a = &r->a[r->order[i]];
      aln *na = (i + 1 == kept) ? NULL : &r->a[r->order[i + 1]];
      char *cg = r->cig + (size_t)a->cigar_buf * r->cig_stride;
      ++a->n_matches;
      a->n_unique += a->length;
      if (a->prefix)
        cig_end += (uint64_t)sprintf(cg + cig_end, "%lu%c", (unsigned long)(a->prefix - last_end), last_end ? 'M' : 'S');
      cig_end += (uint64_t)sprintf(cg + cig_end, "%lu=", (unsigned long)a->length);
      if (!na || na->pos != a->pos || na->seq_index != a->seq_index || na->rc != a->rc) {
        if (a->suffix) cig_end += (uint64_t)sprintf(cg + cig_end, "%luS", (unsigned long)a->suffix);
        for (uint64_t j = 0; j < q; ++j) {                    /* XE, query.cpp:270-274 */
          int64_t rp = a->rcpos + (int64_t)j;
          if (rp >= 0 && rp < (int64_t)ix->N && ix->text[rp] == r->lower[j]) ++a->n_matched;
        }
        cg[cig_end] = 0;
        cig_end = 0; last_end = 0;
      } else {
        last_end = a->prefix + a->length;
        int t = a->cigar_buf; a->cigar_buf = na->cigar_buf; na->cigar_buf = t;   /* cigar.swap */
        na->qpos = a->qpos < na->qpos ? a->qpos : na->qpos;
        uint64_t u = a->n_matches; a->n_matches = na->n_matches; na->n_matches = u;
        a->n_matches = 0;
        u = a->n_unique; a->n_unique = na->n_unique; na->n_unique = u;
        a->n_matched = 0;
      }
    }
    std_sort(r->order, kept, lt_print, r->a);
    r->best = r->order[0];
    int prev = -1;
    for (int i = 0; i < kept; ++i) {
      aln *a = &r->a[r->order[i]];
      if (a->n_matches) {
        a->hi_index = (uint64_t)r->n_aln++;
        if (prev >= 0) { a->prev = prev; r->a[prev].next = r->order[i]; }
        prev = r->order[i];
      }
    }
  }
  if (r->n_aln == 0 && p->nomap) {                            /* set_nomap, query.cpp:308-320 */
    r->n_aln = 1;
    r->read_flag |= 4;
    memset(&r->a[0], 0, sizeof(aln));
    r->a[0].prev = r->a[0].next = -1;
    r->a[0].cigar_buf = 0;
    r->cig[0] = '*'; r->cig[1] = 0;
    r->order[0] = 0;
    r->n = 1;
    r->unmapped_placeholder = 1;
  }
}

typedef struct { char *p; uint64_t len, cap; } obuf;
static void ob_need(obuf *o, uint64_t extra) {
  if (o->len + extra > o->cap) {
    o->cap = (o->len + extra) * 2 + 4096;
    o->p = (char *)realloc(o->p, o->cap);
  }
}
static void ob_put(obuf *o, const void *s, uint64_t n) { ob_need(o, n); memcpy(o->p + o->len, s, n); o->len += n; }
static void ob_str(obuf *o, const char *s) { ob_put(o, s, strlen(s)); }
static void ob_i64(obuf *o, int64_t v) { char t[32]; int n = sprintf(t, "%ld", (long)v); ob_put(o, t, (uint64_t)n); }
static void ob_u64(obuf *o, uint64_t v) { char t[32]; int n = sprintf(t, "%lu", (unsigned long)v); ob_put(o, t, (uint64_t)n); }
static void ob_ch(obuf *o, char c) { ob_put(o, &c, 1); }

static uint8_t comp_char(uint8_t c) {                         /* reverse_complement, fasta.cpp:26-61 */
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

/* mate view of a read for set_mate/print (query.cpp:417-434) */
typedef struct { int has; uint64_t seq_index; int64_t pos; } mate_ref;

/* print_matches (query.cpp:331-415), SAM branch */
static void print_read(const orc_index *ix, const read_state *r, const mate_ref *mate,
                       const uint8_t *name, uint64_t name_len, const uint8_t *seq,
                       const uint8_t *qual, uint64_t q, const uint8_t *opt, uint64_t opt_len,
                       obuf *o) {
  for (int i = 0; i < r->n; ++i) {
    const aln *a = &r->a[r->order[i]];
    const int unmapped = (r->read_flag & 4) != 0;
    if (!(a->n_matches || unmapped)) continue;
    ob_put(o, name, name_len); ob_ch(o, '\t');
    if (unmapped) {
      ob_u64(o, r->read_flag); ob_ch(o, '\t');
      if (mate->has) { ob_str(o, ix->descr[mate->seq_index]); ob_ch(o, '\t'); ob_i64(o, mate->pos + 1); }
      else ob_str(o, "*\t0");
      ob_str(o, "\t0\t*");
    } else {
      ob_u64(o, r->read_flag | (a->rc ? 16u : 0u) | (a->hi_index ? 256u : 0u)); ob_ch(o, '\t');
      ob_str(o, ix->descr[a->seq_index]); ob_ch(o, '\t'); ob_i64(o, a->pos + 1);
      ob_str(o, "\t50\t"); ob_str(o, r->cig + (size_t)a->cigar_buf * r->cig_stride);
    }
    if (mate->has) { ob_ch(o, '\t'); ob_str(o, ix->descr[mate->seq_index]); ob_ch(o, '\t'); ob_i64(o, mate->pos + 1); ob_str(o, "\t0"); }
    else ob_str(o, "\t*\t0\t0");
    ob_ch(o, '\t');
    ob_need(o, 2 * q + 2);
    if (a->rc) {
      for (uint64_t j = 0; j < q; ++j) o->p[o->len + j] = (char)comp_char(seq[q - 1 - j]);
      o->len += q; ob_ch(o, '\t');
      for (uint64_t j = 0; j < q; ++j) o->p[o->len + j] = (char)qual[q - 1 - j];
      o->len += q;
    } else {
      ob_put(o, seq, q); ob_ch(o, '\t'); ob_put(o, qual, q);
    }
    if (a->n_matches) {
      ob_str(o, "\tXM:i:"); ob_u64(o, a->n_matches);
      ob_str(o, "\tXU:i:"); ob_u64(o, a->n_unique);
      ob_str(o, "\tXE:i:"); ob_u64(o, a->n_matched);
      ob_str(o, "\tXS:A:"); ob_ch(o, a->rc ? '-' : '+');
      ob_str(o, "\tNH:i:"); ob_u64(o, (uint64_t)r->n_aln);
      ob_str(o, "\tHI:i:"); ob_u64(o, a->hi_index);
    } else {
      ob_str(o, "\tXM:i:0\tNH:i:0");
    }
    if (a->prev >= 0) {
      const aln *pv = &r->a[a->prev];
      ob_str(o, "\tcc:Z:"); ob_str(o, ix->descr[pv->seq_index]);
      ob_str(o, "\tcp:i:"); ob_i64(o, pv->pos + 1);
      ob_str(o, "\txo:A:"); ob_ch(o, pv->rc == a->rc ? '=' : '!');
      ob_str(o, "\txc:Z:"); ob_str(o, r->cig + (size_t)pv->cigar_buf * r->cig_stride);
    }
    if (a->next >= 0) {
      const aln *nx = &r->a[a->next];
      ob_str(o, "\tCC:Z:"); ob_str(o, ix->descr[nx->seq_index]);
      ob_str(o, "\tCP:i:"); ob_i64(o, nx->pos + 1);
      ob_str(o, "\tXO:A:"); ob_ch(o, nx->rc == a->rc ? '=' : '!');
      ob_str(o, "\tXC:Z:"); ob_str(o, r->cig + (size_t)nx->cigar_buf * r->cig_stride);
    }
    if (opt_len) ob_put(o, opt, opt_len);
    ob_ch(o, '\n');
  }
}

typedef struct {
  const orc_index *ix; const orc_batch *b; const orc_params *p;
  uint64_t first, last;              /* read range [first,last), first even */
  obuf out;
  orc_match *matches; uint64_t n_matches, cap_matches;
  uint64_t *per_read;                /* match count per read (global array) */
} job;

static void job_keep_matches(job *J, const read_state *r, uint64_t n, uint64_t read) {
  if (!J->per_read) return;
  J->per_read[read] = n;
  if (J->n_matches + n > J->cap_matches) {
    J->cap_matches = (J->n_matches + n) * 2 + 1024;
    J->matches = (orc_match *)realloc(J->matches, sizeof(orc_match) * J->cap_matches);
  }
  memcpy(J->matches + J->n_matches, r->m, sizeof(orc_match) * n);
  J->n_matches += n;
}

/* Pair::run (query.cpp:481-520): two reads at a time, mates resolved, then printed. */
static void *job_main(void *arg) {
  job *J = (job *)arg;
  const orc_batch *b = J->b;
  read_state R[2]; memset(R, 0, sizeof R);
  for (int k = 0; k < 2; ++k) { R[k].m_cap = 64; R[k].m = (orc_match *)malloc(sizeof(orc_match) * 64); }
  for (uint64_t i = J->first; i < J->last; i += 2) {
    const int two = (i + 1 < J->last);
    for (int k = 0; k <= two; ++k) {
      uint64_t rd = i + (uint64_t)k;
      uint64_t q = (uint64_t)(b->seq_off[rd + 1] - b->seq_off[rd]);
      run_read(J->ix, J->p, b->seq + b->seq_off[rd], q, b->read_flag[rd], &R[k]);
      job_keep_matches(J, &R[k], R[k].n_found, rd);
    }
    mate_ref mt[2] = {{0, 0, 0}, {0, 0, 0}};
    if (two && (R[0].read_flag & 64) && (R[1].read_flag & 128)) {     /* has_mate + set_mate */
      for (int k = 0; k < 2; ++k) {
        read_state *me = &R[k], *ot = &R[1 - k];
        if (me->n_aln && ot->n_aln) {
          if (ot->best >= 0) { mt[k].has = 1; mt[k].seq_index = ot->a[ot->best].seq_index; mt[k].pos = ot->a[ot->best].pos; }
          else {
            me->read_flag |= 8;
            if (me->best >= 0) { mt[k].has = 1; mt[k].seq_index = me->a[me->best].seq_index; mt[k].pos = me->a[me->best].pos; }
          }
        }
      }
    }
    for (int k = 0; k <= two; ++k) {
      uint64_t rd = i + (uint64_t)k;
      uint64_t q = (uint64_t)(b->seq_off[rd + 1] - b->seq_off[rd]);
      print_read(J->ix, &R[k], &mt[k], b->names + b->name_off[rd],
                 (uint64_t)(b->name_off[rd + 1] - b->name_off[rd]), b->seq + b->seq_off[rd],
                 b->qual + b->seq_off[rd], q, b->opt ? b->opt + b->opt_off[rd] : NULL,
                 b->opt ? (uint64_t)(b->opt_off[rd + 1] - b->opt_off[rd]) : 0, &J->out);
    }
  }
  for (int k = 0; k < 2; ++k) { free(R[k].a); free(R[k].order); free(R[k].cig); free(R[k].m); free(R[k].lower); }
  return NULL;
}

uint64_t orc_map_batch(const orc_index *ix, const orc_batch *b, const orc_params *p, char *out,
                       uint64_t cap, int64_t *match_off, orc_match *matches, uint64_t match_cap) {
  int nt = p->n_threads < 1 ? 1 : p->n_threads;
  uint64_t pairs = (b->n_reads + 1) / 2;
  if ((uint64_t)nt > pairs) nt = pairs ? (int)pairs : 1;
  job *J = (job *)calloc((size_t)nt, sizeof(job));
  pthread_t *th = (pthread_t *)calloc((size_t)nt, sizeof(pthread_t));
  uint64_t *per_read = match_off ? (uint64_t *)calloc(b->n_reads + 1, sizeof(uint64_t)) : NULL;
  for (int t = 0; t < nt; ++t) {
    J[t].ix = ix; J[t].b = b; J[t].p = p; J[t].per_read = per_read;
    J[t].first = 2 * (pairs * (uint64_t)t / (uint64_t)nt);
    J[t].last = 2 * (pairs * (uint64_t)(t + 1) / (uint64_t)nt);
    if (J[t].last > b->n_reads) J[t].last = b->n_reads;
    pthread_create(&th[t], NULL, job_main, &J[t]);
  }
  uint64_t total = 0, mtotal = 0;
  for (int t = 0; t < nt; ++t) {
    pthread_join(th[t], NULL);
    if (total + J[t].out.len <= cap && out) memcpy(out + total, J[t].out.p, J[t].out.len);
    total += J[t].out.len;
    if (match_off) {
      if (matches && mtotal + J[t].n_matches <= match_cap)
        memcpy(matches + mtotal, J[t].matches, sizeof(orc_match) * J[t].n_matches);
      mtotal += J[t].n_matches;
    }
    free(J[t].out.p); free(J[t].matches);
  }
  if (match_off) {
    int64_t acc = 0;
    for (uint64_t i = 0; i < b->n_reads; ++i) { match_off[i] = acc; acc += (int64_t)per_read[i]; }
    match_off[b->n_reads] = acc;
    free(per_read);
  }
  free(J); free(th);
  return total;
}

/* ------------------------------------------------------------------ index build (small) */

typedef struct { const uint8_t *t; uint64_t n; } sufctx;
static int suf_cmp(const void *x, const void *y, void *c) {
  const sufctx *S = (const sufctx *)c;
  uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
  uint64_t la = S->n - a, lb = S->n - b, l = la < lb ? la : lb;
  int r = memcmp(S->t + a, S->t + b, l);
  if (r) return r;
  return la < lb ? -1 : (la > lb);
}
/* SA/ISA/LCP are canonical functions of the text (qsufsort.cpp:266-344 + computeLCP,
 * longSA.cpp:224-237 produce the same arrays), so any correct builder is a valid oracle. */
void orc_build_index(const uint8_t *text, uint64_t N, uint64_t *sa, uint64_t *isa, uint64_t *lcp) {
  for (uint64_t i = 0; i < N; ++i) sa[i] = i;
  sufctx S = {text, N};
  qsort_r(sa, N, sizeof(uint64_t), suf_cmp, &S);
  for (uint64_t i = 0; i < N; ++i) isa[sa[i]] = i;
  uint64_t h = 0;
  for (uint64_t i = 0; i < N; ++i) {
    uint64_t m = isa[i];
    if (m == 0) { lcp[m] = 0; }
    else {
      uint64_t j = sa[m - 1];
      while (i + h < N && j + h < N && text[i + h] == text[j + h]) ++h;
      lcp[m] = h;
    }
    if (h) --h;
  }
}

/* longSA::show, bin=true (longSA.cpp:612-690) */
void orc_mappability(const orc_index *ix, uint8_t *out) {
  const uint64_t N = ix->N;
  uint64_t *ml = (uint64_t *)calloc(N, sizeof(uint64_t));
  for (uint64_t i = 0; i < N; ++i) {
    ml[i] = lcp_at(ix, i) + 1;
    if (i && ml[i] > ml[i - 1]) ml[i - 1] = ml[i];
  }
  uint64_t w = 0;
  for (uint64_t c = 0; c < ix->n_descr; c += 2) {
    const uint64_t start = ix->startpos[c], size = ix->sizes[c];
    for (uint64_t i = 0; i < size; ++i) {
      const uint64_t sp = isa_at(ix, start + i);
      const uint64_t rp = isa_at(ix, start + 2 * size - i);
      if (ml[sp] + i >= size) ml[sp] = 0;
      if (ml[rp] >= i) ml[rp] = 0;
      out[w++] = (uint8_t)(ml[rp] < 255 ? ml[rp] : 255);
      out[w++] = (uint8_t)(ml[sp] < 255 ? ml[sp] : 255);
    }
  }
  free(ml);
}
