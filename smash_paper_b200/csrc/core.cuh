// core.cuh -- data layout in HBM and the per-lane building blocks of the kernels.
//
// Everything here is `__host__ __device__` so that tests/emul (a host-side lane-by-lane
// emulation used only by the CPU test-suite to debug logic without a GPU) can run exactly the
// same functions the kernels run.  The product never executes these on the host.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#if defined(SMASH_EXACT_INLINE)
#define HDNI __host__ __device__
#else
#define HDNI __host__ __device__ __noinline__
#endif
#else
#define HD inline
#define HDN
#define HDNI
#endif

namespace smash {

constexpr int TEXT_PAD = 64;        // zero bytes before text[0] and after text[N-1]
constexpr int MAXQ_FAST = 1024;     // reads longer than this take the global-memory path
constexpr int P_FRONT = 16;         // bytes of 0xFE before the staged read
constexpr int P_BACK = 16;          // bytes of 0xFF after it
constexpr int BIG_BUCKET = 48;      // seed buckets larger than this use the exact per-start search
constexpr int STAGE_CAP = 64;       // matches staged in shared memory per read before spilling

struct LcpItem { uint64_t idx, val; };          // .lcp.m.bin item (longSA.h:19-28), val widened

// The reference's index as it lives in HBM (+ the structures derived from it at load time).
struct DevIndex {
  const uint8_t *text;       // N bytes (rc1.ref.seq.bin), TEXT_PAD zero bytes either side
  uint64_t N;
  const void *sa;            // N * w
  const void *isa;           // N * w or null (only MEM mode / mappability build need it)
  int w;                     // 4 or 8 (size.h:9-22)
  const uint8_t *lcp;        // N bytes, 255 => look in lcp_m (longSA.h:34-39)
  const LcpItem *lcp_m;
  uint64_t n_m;
  // derived at load (file surface unchanged):
  const uint8_t *uniq;       // U[pos] = min(255, max(LCP[ISA[pos]], LCP[ISA[pos]+1]) + 1): shortest unique length
  const void *seed;          // S[x] = #suffixes < k-mer x (monotone, 4^k + 1 entries)
  int seed_k;
  int seed_w;                // 4 or 8 bytes per entry
#ifndef SMASH_SEED_FIELDS_LAST
  // Blocked form of the 8-byte seed table (seed_blocked != 0; built in place, same bytes): 16 consecutive k-mer buckets
  // share one 128-byte line that also holds the ext codes of the block's first SEED_INLINE suffixes, so an anchor's
  // bucket bounds AND its candidates' pre-filter codes arrive with ONE line fill from HBM instead of two (seed_block_*).
  int seed_blocked;
  uint64_t seed_n;           // 4^seed_k buckets
  uint64_t seed_end;         // S[4^seed_k]
  const uint64_t *seed_irr;  // exact S[16B .. 16B+16] of the blocks holding a bucket of >= 255 suffixes
#endif
  const uint32_t *ext;       // per SA rank i: 2-bit codes of the 6 text chars after SA[i]+k, the 8 before SA[i], and how many of those are acgt (ext_entry)
  uint32_t alpha[8];         // bitmap of byte values present in text
  // Sequence metadata (fasta.h:24-42)
  const uint64_t *startpos;  // n_descr
  const uint64_t *sizes;
  int n_descr;
  int rcref;
  const char *descr;         // concatenated names
  const int *descr_off;      // n_descr + 1
  const uint64_t *descr8;    // first 8 name bytes packed little-endian (used when the name has <= 8 chars)
  uint64_t logN;             // ceil(log2 N) as longSA.cpp:97 computes it
  // mappability (map.bin body) + 32-bit chromosome offsets of mappability_tag (chromosomes.h:109)
  const uint8_t *mapbody;
  uint64_t map_bytes;
  const uint32_t *chrom_off32;   // per forward chromosome
  const uint64_t *chrom_abs64;   // MemSam::chromosomes (query.cpp:546-552): 64-bit running sums per forward chromosome, [n_fwd] = total ("*")
#ifdef SMASH_SEED_FIELDS_LAST
  // Blocked form of the 8-byte seed table (seed_blocked != 0; built in place, same bytes): 16 consecutive k-mer buckets
  // share one 128-byte line that also holds the ext codes of the block's first SEED_INLINE suffixes, so an anchor's
  // bucket bounds AND its candidates' pre-filter codes arrive with ONE line fill from HBM instead of two (seed_block_*).
  int seed_blocked;
  uint64_t seed_n;           // 4^seed_k buckets
  uint64_t seed_end;         // S[4^seed_k]
  const uint64_t *seed_irr;  // exact S[16B .. 16B+16] of the blocks holding a bucket of >= 255 suffixes
#endif
};

HD uint64_t sa_at(const DevIndex &ix, uint64_t i) {
  return ix.w == 4 ? (uint64_t)((const uint32_t *)ix.sa)[i] : ((const uint64_t *)ix.sa)[i];
}
HD uint64_t isa_at(const DevIndex &ix, uint64_t i) {
  return ix.w == 4 ? (uint64_t)((const uint32_t *)ix.isa)[i] : ((const uint64_t *)ix.isa)[i];
}
// ---- blocked seed table.  Block B = buckets 16B .. 16B+15, one 128-byte line of 16 words:
//   word 0        bits 0..39 S[16B] (rank of the block's first suffix), bit 63 = irregular;
//   words 1..2    the 16 bucket sizes S[x+1]-S[x] as bytes (regular blocks: all < 255); irregular: word 1 = row in seed_irr;
//   bytes 24..127 ext[S[16B] + j], j < SEED_INLINE: the pre-filter codes of the block's first 26 suffixes (the suffix array
//                 is in bucket order, so these are the block's buckets' candidates; later ones stay in the flat ext array).
// A random 3.1 Gb genome with k = 16 has 23 suffixes per block: ~96 % of the anchors find all they need in the one line.
constexpr int SEED_INLINE = 26;
constexpr uint64_t SEED_BASE_MASK = (1ull << 40) - 1ull;
HD uint32_t bytesum8(uint64_t v) {                       // sum of the 8 bytes of v
  const uint64_t m = 0x00ff00ff00ff00ffull;
  return (uint32_t)((((v & m) + ((v >> 8) & m)) * 0x0001000100010001ull) >> 48);
}
// sum of the first i (0..15) count bytes of a regular block
HD uint32_t seed_block_prefix(uint64_t c0, uint64_t c1, unsigned i) {
  if (i < 8) return i ? bytesum8(c0 & ((1ull << (8 * i)) - 1ull)) : 0u;
  return bytesum8(c0) + (i > 8 ? bytesum8(c1 & ((1ull << (8 * (i - 8))) - 1ull)) : 0u);
}
// S[x] from the blocked table: the exact paths' lookup (kept out of line: the hot loops around it must not grow)
HDNI inline uint64_t seed_at_blocked(const DevIndex &ix, uint64_t x) {
  if (x >= ix.seed_n) return ix.seed_end;
  const uint64_t *blk = (const uint64_t *)ix.seed + ((x >> 4) << 4);
  const uint64_t h = blk[0];
  const unsigned i = (unsigned)(x & 15);
  if (h >> 63) return ix.seed_irr[blk[1] * 17 + i];
  return (h & SEED_BASE_MASK) + seed_block_prefix(blk[1], blk[2], i);
}
HD uint64_t seed_at(const DevIndex &ix, uint64_t x) {
#ifndef SMASH_NO_BLOCKED
  if (ix.seed_blocked) return seed_at_blocked(ix, x);
#endif
  return ix.seed_w == 4 ? (uint64_t)((const uint32_t *)ix.seed)[x] : ((const uint64_t *)ix.seed)[x];
}
// vec_uchar::operator[] (longSA.h:34-39)
HD uint64_t lcp_at(const DevIndex &ix, uint64_t i) {
  uint8_t v = ix.lcp[i];
  if (v != 255) return v;
  uint64_t lo = 0, hi = ix.n_m;
  while (lo < hi) {
    uint64_t mid = lo + ((hi - lo) >> 1);
    if (ix.lcp_m[mid].idx < i) lo = mid + 1; else hi = mid;
  }
  return ix.lcp_m[lo].val;
}

// Unaligned little-endian 8-byte load from the padded text (pos may be in [-TEXT_PAD, N+TEXT_PAD-8]).
HD uint64_t text8(const uint8_t *T, int64_t pos) {
#if defined(__CUDA_ARCH__)
  const uint64_t *a = reinterpret_cast<const uint64_t *>(T + (pos & ~(int64_t)7));
  const unsigned sh = (unsigned)(pos & 7) * 8u;
  uint64_t lo = __ldg(a);
  if (sh == 0) return lo;
  uint64_t hi = __ldg(a + 1);
  return (lo >> sh) | (hi << (64u - sh));
#else
  uint64_t v; memcpy(&v, T + pos, 8); return v;
#endif
}
// Same for the staged read (shared memory on the device). P points at read[0]; the buffer has
// P_FRONT bytes before and P_BACK after, so pos in [-8, q+8) is fine.
HD uint64_t read8(const uint8_t *P, int pos) {
#if defined(__CUDA_ARCH__)
  const uint8_t *b = P + pos;
  const uint32_t *a = reinterpret_cast<const uint32_t *>(b - ((uintptr_t)b & 3));
  const unsigned sh = (unsigned)((uintptr_t)b & 3) * 8u;
  uint32_t w0 = a[0], w1 = a[1], w2 = a[2];
  uint64_t lo = ((uint64_t)w1 << 32) | w0;
  if (sh == 0) return lo;
  return (lo >> sh) | ((uint64_t)w2 << (64u - sh));
#else
  uint64_t v; memcpy(&v, P + pos, 8); return v;
#endif
}

HD int ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}

// a,c,g,t -> 0..3 (lexicographic), anything else -> 4
HD int base_code(uint8_t c) {
  return c == 'a' ? 0 : c == 'c' ? 1 : c == 'g' ? 2 : c == 't' ? 3 : 4;
}
HD bool in_alpha(const DevIndex &ix, uint8_t c) { return (ix.alpha[c >> 5] >> (c & 31)) & 1u; }

// NewQuery::extend (query.cpp:125-144): tolower, and with -n everything but acgt becomes '~'.
HD uint8_t query_char(uint8_t c, int nucleotides_only) {
  if (c >= 'A' && c <= 'Z') c = (uint8_t)(c + 32);
  if (nucleotides_only && c != 'a' && c != 'c' && c != 'g' && c != 't') c = '~';
  return c;
}

// Number of leading equal bytes of text[tpos..] and read[ppos..], at most `limit`.
HD int match_right(const uint8_t *T, int64_t tpos, const uint8_t *P, int ppos, int limit) {
  int m = 0;
  while (m < limit) {
    uint64_t d = text8(T, tpos + m) ^ read8(P, ppos + m);
    if (d) { m += ctz64(d) >> 3; break; }
    m += 8;
  }
  return m < limit ? m : limit;
}
// Number of equal bytes going left from text[tpos-1], read[ppos-1], at most `limit`.
HD int match_left(const uint8_t *T, int64_t tpos, const uint8_t *P, int ppos, int limit) {
  int m = 0;
  while (m < limit) {
    uint64_t d = text8(T, tpos - m - 8) ^ read8(P, ppos - m - 8);
    if (d) { m += clz64(d) >> 3; break; }
    m += 8;
  }
  return m < limit ? m : limit;
}

struct Match { uint64_t ref; uint32_t qpos, len; };     // 16 bytes in HBM

// Is text[r .. r+len) unique in the text?  U answers in one byte unless it saturated.
HD bool is_unique(const DevIndex &ix, uint64_t r, uint32_t len, uint64_t sa_index_if_known,
                  bool have_sa_index) {
  uint8_t u = ix.uniq[r];
  if (u != 255) return len >= u;
  if (len < 255) return false;
  // saturated: need the exact LCP values around the suffix (overflow table, longSA.h:34-39)
  uint64_t i;
  if (have_sa_index) i = sa_index_if_known;
  else if (ix.isa) i = isa_at(ix, r);
  else return false;   // caller must supply the SA index when ISA is absent
  uint64_t a = lcp_at(ix, i);
  uint64_t b = (i + 1 < ix.N) ? lcp_at(ix, i + 1) : 0;
  return a < len && b < len;
}

// Exact answer for ONE start p: longest prefix of P[p..q) in the text by
// binary search over the suffix array (optionally inside the seed bucket), its uniqueness and
// left-maximality.  Returns true and fills m when (p) is a reportable MAM.
// (kept out of line on the device: it is the rare path and is called from three places of k_mam_search)
// SEEDED = false: the search starts from the whole suffix array (a few more steps).  k_mam_verify, which gets here only for
// saturated repeat families, uses that form: with the seed-table lookup compiled into its call tree ptxas schedules its
// hot loop 20 % slower (measured, profiles/r02 A/B notes in DESIGN.md).
template <bool SEEDED = true>
HDNI inline bool exact_start(const DevIndex &ix, const uint8_t *P, int q, int p, uint32_t L, Match *m) {
  if (q - p < (int)L) return false;
  const uint8_t *T = ix.text;
  uint64_t lo = 0, hi = ix.N;
  int kk = ix.seed_k < (int)L ? ix.seed_k : (int)L;
  if (kk > q - p) kk = q - p;
  if (SEEDED && kk > 0 && ix.seed) {
    uint64_t code = 0; bool ok = true;
    for (int j = 0; j < kk; ++j) { int b = base_code(P[p + j]); if (b > 3) { ok = false; break; } code = (code << 2) | (uint64_t)b; }
    if (ok) {
      int sh = 2 * (ix.seed_k - kk);
      lo = seed_at(ix, code << sh);
      hi = seed_at(ix, (code + 1) << sh);
    }
  }
  if (lo >= hi) return false;
  // insertion point of P[p..] in [lo,hi): l/r are exclusive virtual bounds with lcp 0
  int64_t l = (int64_t)lo - 1, r = (int64_t)hi;
  int llcp = 0, rlcp = 0;
  const int rem = q - p;
  while (r - l > 1) {
    int64_t mid = l + ((r - l) >> 1);
    uint64_t s = sa_at(ix, (uint64_t)mid);
    int h = llcp < rlcp ? llcp : rlcp;
    int lc = h + match_right(T, (int64_t)s + h, P, p + h, rem - h);
    bool greater;   // P[p..] > suffix ?
    if (lc == rem) greater = false;
    else greater = P[p + lc] > ((s + (uint64_t)lc < ix.N) ? T[s + lc] : 0);
    if (greater) { l = mid; llcp = lc; } else { r = mid; rlcp = lc; }
  }
  int best = llcp > rlcp ? llcp : rlcp;
  if (best < (int)L || best < 2) return false;
  if (llcp == rlcp) return false;                    // two suffixes share the longest match
  uint64_t idx = (uint64_t)(llcp > rlcp ? l : r);
  uint64_t ref = sa_at(ix, idx);
  if (!is_unique(ix, ref, (uint32_t)best, idx, true)) return false;
  if (!(p == 0 || ref == 0 || P[p - 1] != T[ref - 1])) return false;   // is_leftmaximal, longSA.cpp:540-546
  m->ref = ref; m->qpos = (uint32_t)p; m->len = (uint32_t)best;
  return true;
}

// ---- anchor path building blocks ------------------------------------------------
// 2-bit codes (a,c,g,t -> 0..3) of 4 consecutive read bytes, first byte in the top bits of the result
// byte.  Bytes must be acgt (validity is tracked separately in a per-read bit mask).
HD uint32_t code4(uint32_t w) {
  uint32_t y = (w >> 1) & 0x03030303u;          // a,c,g,t -> 0,1,3,2
  y ^= (y >> 1) & 0x01010101u;                  //         -> 0,1,2,3
  return (y * 0x40100401u) >> 24;               // c0<<6 | c1<<4 | c2<<2 | c3
}
// k-mer code of P[x..x+k), k <= 16
HD uint32_t kmer_code(const uint8_t *P, int x, int k) {
  const uint64_t w0 = read8(P, x), w1 = read8(P, x + 8);
  const uint32_t c = (code4((uint32_t)w0) << 24) | (code4((uint32_t)(w0 >> 32)) << 16) |
                     (code4((uint32_t)w1) << 8) | code4((uint32_t)(w1 >> 32));
  return k >= 16 ? c : c >> (2 * (16 - k));
}
// does P[x..x+k) contain a non-acgt byte?  inv: one bit per base (bit j of word j>>5), one spare word
HD bool kmer_invalid(const uint32_t *inv, int x, int k) {
  const int wd = x >> 5, sh = x & 31;
  uint32_t m = inv[wd] >> sh;
  if (sh) m |= inv[wd + 1] << (32 - sh);
  return (m & ((k >= 32 ? 0u : (1u << k)) - 1u)) != 0;
}
// ---- 8+6 character pre-filter ------------------------------------------------------------------------
// ext[i] (32 bits) describes the text around the suffix c = SA[i] in 2-bit codes (a,c,g,t -> 0..3, anything else -> 0):
//   bits  0..11  the 6 characters after the k-mer, T[c+k .. c+k+6), first one in the top bits;
//   bits 12..27  the 8 characters before it, T[c-8 .. c), in text order (T[c-1] in bits 13:12);
//   bits 28..31  lv: how many of the characters immediately before c are acgt (0..8; stops at the first other byte or at
//                the start of the text).
// One coalesced 4-byte load per bucket entry therefore tells, before the SA entry or the text is touched,
//   * whether the candidate can reach length L at all (most chance hits of the seed end here), and
//   * its LEFT extension exactly whenever that is shorter than 8 -- i.e. whether an earlier anchor owns this diagonal
//     (left >= stride) -- because within the lv valid characters equal codes are equal bytes, and beyond them the text
//     holds a byte no anchor-path read can match.
// The right-hand codes are only an upper bound (codes of other bytes may agree by accident): a pass is verified on the text.
HD uint32_t ext_entry(const uint8_t *T, uint64_t N, uint64_t c, int k) {
  uint32_t r = 0, l = 0, lv = 0; bool open = true;
  for (int j = 0; j < 6; ++j) {
    const uint64_t pr = c + (uint64_t)k + (uint64_t)j;
    const int br = pr < N ? base_code(T[pr]) : 4;
    r = (r << 2) | (uint32_t)(br > 3 ? 0 : br);
  }
  for (int j = 1; j <= 8; ++j) {                               // T[c-1], T[c-2], ..
    const int bl = c >= (uint64_t)j ? base_code(T[c - (uint64_t)j]) : 4;
    l |= (uint32_t)(bl > 3 ? 0 : bl) << (2 * (j - 1));
    if (bl > 3) open = false;
    if (open) ++lv;
  }
  return r | (l << 12) | (lv << 28);
}
HD uint32_t ext_right_pairs(uint32_t x12) {   // leading equal 2-bit groups of a 12-bit xor (0..6)
#if defined(__CUDA_ARCH__)
  return x12 ? (uint32_t)(__clz((int)x12) - 20) >> 1 : 6u;
#else
  return x12 ? (uint32_t)(__builtin_clz(x12) - 20) >> 1 : 6u;
#endif
}
HD uint32_t ext_left_pairs(uint32_t x16) {    // trailing equal 2-bit groups of a 16-bit xor (0..8)
#if defined(__CUDA_ARCH__)
  return x16 ? (uint32_t)(__ffs((int)x16) - 1) >> 1 : 8u;
#else
  return x16 ? (uint32_t)__builtin_ctz(x16) >> 1 : 8u;
#endif
}
// the read's side of the comparison, from the staged bytes: P[x-8 .. x) and P[x+k .. x+k+6) (lv bits zero)
HD uint32_t read_ext_codes(const uint8_t *P, int x, int k) {
  const uint64_t r = read8(P, x + k), l = read8(P, x - 8);
  const uint32_t r12 = (code4((uint32_t)r) << 4) | (code4((uint32_t)(r >> 32)) >> 4);
  const uint32_t l16 = (code4((uint32_t)l) << 8) | code4((uint32_t)(l >> 32));      // P[x-8] in the top bits
  return r12 | (l16 << 12);
}
// conservative form (no validity information about the read): can a candidate with these codes still reach length L?
HD bool ext_may_reach(uint32_t cand_ext, uint32_t read_ext, int k, uint32_t L) {
  const uint32_t x = cand_ext ^ read_ext;
  const uint32_t r = ext_right_pairs(x & 0xfffu), l = ext_left_pairs((x >> 12) & 0xffffu);
  return r == 6 || l == 8 || (uint32_t)k + r + l >= L;
}

// seed bucket of anchor x: [lo, hi) in SA order (a sorted superset of the k-mer's interval)
HD void anchor_bucket(const DevIndex &ix, const uint8_t *P, int x, int k, uint64_t *lo, uint64_t *hi) {
  const uint64_t code = kmer_code(P, x, k);
  const int sh = 2 * (ix.seed_k - k);
  *lo = seed_at(ix, code << sh);
  *hi = seed_at(ix, (code + 1) << sh);
}
// One candidate suffix c = SA[i] of anchor x (stride s): extend along the diagonal.
//   returns 1 and fills m   -> a reportable match owned by this anchor
//           0               -> nothing (junk, owned by an earlier anchor, too short, not unique)
//          -1               -> U saturated: *pl holds the start whose exact answer the caller must compute
HD int candidate_check(const DevIndex &ix, const uint8_t *P, int q, int x, int s, int k, uint32_t L, uint64_t c,
                       Match *m, int *pl) {
  const uint8_t *T = ix.text;
  // ownership first: the 0xFE pad before the read and the zero pad before the text stop this at either start
  const int left = match_left(T, (int64_t)c, P, x, s);
  if (left >= s) return 0;               // start <= x-s: an earlier anchor owns this diagonal
  const int right = match_right(T, (int64_t)c, P, x, q - x);
  if (right < k) return 0;               // bucket is a superset: suffixes between two k-mers
  const uint32_t len = (uint32_t)(left + right);
  if (len < L || len < 2) return 0;
  const uint64_t ref = c - (uint64_t)left;
  const uint8_t u = ix.uniq[ref];
  if (u == 255 && len >= 255) { *pl = x - left; return -1; }
  if (len < u) return 0;
  m->ref = ref; m->qpos = (uint32_t)(x - left); m->len = len;
  return 1;
}

}  // namespace smash
