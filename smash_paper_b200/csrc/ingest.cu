// ingest.cu -- kernels of the device-side input stage (semantics in ingest.cuh).
//
//   raw text in HBM
//     -> line starts        (k_ing_scan_* over 16-byte chunks: count, scan, place)
//     -> [FASTQ] reader state before every line (scan of per-line transition functions under
//        composition), header lines compacted into record -> line tables
//     -> one LineRec per SAM line / FASTQ record (thread per line: sequential tokeniser, bytes stay in L1)
//     -> exclusive scan of {reads, name bytes, seq bytes, opt bytes}  (ordered compaction: output order
//        = input order, mates of a FASTQ pair interleaved as fastqs_to_sam prints them)
//     -> k_ing_copy (8 lanes per read): names / SEQ / QUAL / optional fields into the packed batch blobs as aligned
//        16-byte stores assembled with funnel shifts, offsets and read_flag; whitespace runs of the optional
//        fields collapse to one tab.
// HBM-bound streaming work: the text is read three times (count, place, parse+copy; the last two hit L2
// for batches under ~100 MB) and the batch (~ the same bytes) is written once.
#include <cuda_runtime.h>

#include "ingest_launch.cuh"

namespace smash {

#ifndef ING_IB
#define ING_IB 256
#endif
#ifndef ING_II
#define ING_II 4
#endif
constexpr int IB = ING_IB;         // threads per block (the CPU test-suite also builds a small-tile variant)
constexpr int II = ING_II;         // items per thread
static_assert(IB % 32 == 0 && IB >= 32 && IB <= 1024 && II >= 1, "scan tile geometry");
constexpr int IT = IB * II;        // items per tile

uint64_t ing_tiles(uint64_t n) { return (n + IT - 1) / IT; }

// every launch goes through here (the CPU test-suite substitutes a host executor, tests/emul/cuda_shim)
template <class... KArgs, class... Args>
static void ing_launch(void (*k)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
#if defined(__CUDACC__)
  k<<<grid, block, 0, st>>>(args...);
#else
  (void)st;
  shim_launch(k, grid, block, args...);
#endif
}

// ---- generic tiled scan: V value type, In(i) -> V, Op(a, b) = "a then b" (associative), Out(i, prefix, v) ----
__device__ __forceinline__ uint32_t shfl_up_v(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ uint64_t shfl_up_v(uint64_t v, int d) { return __shfl_up_sync(0xffffffffu, (unsigned long long)v, d); }
__device__ __forceinline__ Ing4 shfl_up_v(Ing4 v, int d) {
  Ing4 r;
  r.reads = shfl_up_v(v.reads, d); r.name = shfl_up_v(v.name, d); r.seq = shfl_up_v(v.seq, d); r.opt = shfl_up_v(v.opt, d);
  return r;
}

struct OpAdd64 { __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a + b; } };
struct OpAdd4 { __device__ __forceinline__ Ing4 operator()(const Ing4 &a, const Ing4 &b) const { return ing4_add(a, b); } };
struct OpFsm { __device__ __forceinline__ uint32_t operator()(uint32_t f, uint32_t g) const { return fq_compose(f, g); } };

// exclusive scan of one value per thread over the block; *total = reduction of the whole block.
// wsum: shared, IB/32 + 1 entries.  All IB threads must call it.
template <class V, class Op>
__device__ V block_scan_excl(V v, Op op, V ident, V *wsum, V *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  V inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const V t = shfl_up_v(inc, o); if (lane >= o) inc = op(t, inc); }
  V ex = shfl_up_v(inc, 1);
  if (lane == 0) ex = ident;
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const V w = lane < IB / 32 ? wsum[lane] : ident;
    V winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const V t = shfl_up_v(winc, o); if (lane >= o) winc = op(t, winc); }
    V wex = shfl_up_v(winc, 1);
    if (lane == 0) wex = ident;
    if (lane < IB / 32) wsum[lane] = wex;
    if (lane == IB / 32 - 1) wsum[IB / 32] = winc;
  }
  __syncthreads();
  const V r = op(wsum[warp], ex);
  *total = wsum[IB / 32];
  __syncthreads();
  return r;
}

template <class V, class In, class Op>
__global__ void __launch_bounds__(IB) k_ing_scan_tiles(In in, uint64_t n, V *__restrict__ blk, Op op, V ident) {
  __shared__ V wsum[IB / 32 + 1];
  const uint64_t base = (uint64_t)blockIdx.x * IT + (uint64_t)threadIdx.x * II;
  V s = ident;
#pragma unroll
  for (int i = 0; i < II; ++i) if (base + i < n) s = op(s, in(base + i));
  V total;
  block_scan_excl(s, op, ident, wsum, &total);
  if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

// blk[0..n_blk) -> exclusive prefixes, blk[n_blk] = grand total (one block, IT entries per round)
template <class V, class Op>
__global__ void __launch_bounds__(IB) k_ing_scan_top(V *blk, uint64_t n_blk, Op op, V ident) {
  __shared__ V wsum[IB / 32 + 1];
  __shared__ V carry;
  if (threadIdx.x == 0) carry = ident;
  __syncthreads();
  for (uint64_t b0 = 0; b0 < n_blk; b0 += IT) {
    const uint64_t base = b0 + (uint64_t)threadIdx.x * II;
    V v[II];
    V s = ident;
#pragma unroll
    for (int i = 0; i < II; ++i) { v[i] = base + i < n_blk ? blk[base + i] : ident; s = op(s, v[i]); }
    V total;
    const V ex = block_scan_excl(s, op, ident, wsum, &total);
    const V c = carry;
    V run = op(c, ex);
#pragma unroll
    for (int i = 0; i < II; ++i) { if (base + i < n_blk) blk[base + i] = run; run = op(run, v[i]); }
    __syncthreads();
    if (threadIdx.x == 0) carry = op(c, total);
    __syncthreads();
  }
  if (threadIdx.x == 0) blk[n_blk] = carry;
}

template <class V, class In, class Op, class Out>
__global__ void __launch_bounds__(IB) k_ing_scan_apply(In in, uint64_t n, const V *__restrict__ blk, Op op, V ident, Out out) {
  __shared__ V wsum[IB / 32 + 1];
  const uint64_t base = (uint64_t)blockIdx.x * IT + (uint64_t)threadIdx.x * II;
  V v[II];
  V s = ident;
#pragma unroll
  for (int i = 0; i < II; ++i) { v[i] = base + i < n ? in(base + i) : ident; s = op(s, v[i]); }
  V total;
  V ex = block_scan_excl(s, op, ident, wsum, &total);
  ex = op(blk[blockIdx.x], ex);
#pragma unroll
  for (int i = 0; i < II; ++i) {
    if (base + i < n) out(base + i, ex, v[i]);
    ex = op(ex, v[i]);
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out.total(n, blk[gridDim.x]);
}

template <class V, class In, class Op, class Out>
static int scan3(In in, uint64_t n, V *blk, Op op, V ident, Out out, cudaStream_t st) {
  const unsigned n_blk = (unsigned)ing_tiles(n);
  ing_launch(k_ing_scan_tiles<V, In, Op>, n_blk, IB, st, in, n, blk, op, ident);
  ing_launch(k_ing_scan_top<V, Op>, 1, IB, st, blk, n_blk, op, ident);
  ing_launch(k_ing_scan_apply<V, In, Op, Out>, n_blk, IB, st, in, n, blk, op, ident, out);
  return 3;
}

// ---- line starts --------------------------------------------------------------------------------
struct ChunkIn {
  const uint8_t *text; uint64_t n;
  __device__ __forceinline__ uint64_t operator()(uint64_t c) const {
    const uint4 v = *reinterpret_cast<const uint4 *>(text + 16 * c);     // buffer is 256-byte aligned and padded
    return ing_chunk_starts((uint64_t)v.x | ((uint64_t)v.y << 32), (uint64_t)v.z | ((uint64_t)v.w << 32), 16 * c, n);
  }
};
struct LineStartOut {
  const uint8_t *text; uint64_t n; uint64_t *ls;
  __device__ __forceinline__ void operator()(uint64_t c, uint64_t ex, uint64_t cnt) const {
    if (!cnt) return;
    const uint4 v = *reinterpret_cast<const uint4 *>(text + 16 * c);
    ing_chunk_place((uint64_t)v.x | ((uint64_t)v.y << 32), (uint64_t)v.z | ((uint64_t)v.w << 32), 16 * c, n, ls, ex);
  }
  __device__ __forceinline__ void total(uint64_t, uint64_t tot) const { ls[tot] = ing_sentinel(text, n); }
};

int launch_ing_count_lines(const uint8_t *text, uint64_t n, uint64_t *blk, cudaStream_t st) {
  const uint64_t n_chunks = (n + 15) / 16;
  const unsigned n_blk = (unsigned)ing_tiles(n_chunks);
  ing_launch(k_ing_scan_tiles<uint64_t, ChunkIn, OpAdd64>, n_blk, IB, st, ChunkIn{text, n}, n_chunks, blk, OpAdd64(), 0ull);
  ing_launch(k_ing_scan_top<uint64_t, OpAdd64>, 1, IB, st, blk, n_blk, OpAdd64(), 0ull);
  return 2;
}
int launch_ing_line_starts(const uint8_t *text, uint64_t n, const uint64_t *blk, uint64_t *ls, cudaStream_t st) {
  const uint64_t n_chunks = (n + 15) / 16;
  const unsigned n_blk = (unsigned)ing_tiles(n_chunks);
  ing_launch(k_ing_scan_apply<uint64_t, ChunkIn, OpAdd64, LineStartOut>, n_blk, IB, st, ChunkIn{text, n}, n_chunks, blk, OpAdd64(), 0ull,
                                                                                   LineStartOut{text, n, ls});
  return 1;
}

// ---- SAM lines ------------------------------------------------------------------------------------
__device__ __forceinline__ void ing_report(unsigned long long *err, uint64_t index, uint32_t code) {
  if (code) atomicMin(err, (unsigned long long)((index << 8) | code));
}

__global__ void __launch_bounds__(128) k_ing_parse_sam(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls, uint64_t n_lines,
                                                       LineRec *__restrict__ recs, unsigned long long *err) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_lines; j += (uint64_t)gridDim.x * blockDim.x) {
    LineRec r;
    ing_parse_sam_line(text, ls[j], ing_line_end(ls, j), r);
    recs[j] = r;
    ing_report(err, j, r.err);
  }
}
int launch_ing_parse_sam(const uint8_t *text, const uint64_t *ls, uint64_t n_lines, LineRec *recs, unsigned long long *err, cudaStream_t st) {
  const uint64_t want = (n_lines + 127) / 128;
  const unsigned grid = (unsigned)(want < 148ull * 64 ? (want ? want : 1) : 148ull * 64);
  ing_launch(k_ing_parse_sam, grid, 128, st, text, ls, n_lines, recs, err);
  return 1;
}

// ---- FASTQ ----------------------------------------------------------------------------------------
struct FsmIn {
  const uint8_t *text; const uint64_t *ls;
  __device__ __forceinline__ uint32_t operator()(uint64_t j) const { return fq_line_fn(ing_first_char(text, ls[j], ing_line_end(ls, j))); }
};
struct FsmOut {                       // header line = non-blank line met in state H
  uint8_t *hdr_flag;
  __device__ __forceinline__ void operator()(uint64_t j, uint32_t before, uint32_t fn) const {
    hdr_flag[j] = (fq_apply(before, FQ_H) == FQ_H && fq_apply(fn, FQ_H) != FQ_H) ? 1 : 0;
  }
  __device__ __forceinline__ void total(uint64_t, uint32_t) const {}
};
struct FlagIn {
  const uint8_t *flag;
  __device__ __forceinline__ uint64_t operator()(uint64_t j) const { return flag[j]; }
};
struct HdrOut {                       // record k of the file -> its header line; *count = records in the file
  uint64_t *hdr; uint64_t *count;
  __device__ __forceinline__ void operator()(uint64_t j, uint64_t k, uint64_t is_hdr) const { if (is_hdr) hdr[k] = j; }
  __device__ __forceinline__ void total(uint64_t, uint64_t tot) const { *count = tot; }
};
int launch_ing_fastq_headers(const uint8_t *text, const uint64_t *ls, uint64_t n_lines, uint32_t *blk32, uint64_t *blk64, uint8_t *hdr_flag,
                             uint64_t *hdr, uint64_t *count, cudaStream_t st) {
  int nl = scan3<uint32_t>(FsmIn{text, ls}, n_lines, blk32, OpFsm(), FQ_IDENT, FsmOut{hdr_flag}, st);
  nl += scan3<uint64_t>(FlagIn{hdr_flag}, n_lines, blk64, OpAdd64(), 0ull, HdrOut{hdr, count}, st);
  return nl;
}

__global__ void __launch_bounds__(128) k_ing_parse_fastq(const uint8_t *__restrict__ text, const uint64_t *__restrict__ ls, uint64_t n_lines,
                                                         const uint64_t *__restrict__ hdr, uint64_t n_take, int file, int phase, int replace_n,
                                                         LineRec *__restrict__ recs, unsigned long long *err) {
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_take; k += (uint64_t)gridDim.x * blockDim.x) {
    LineRec r;
    ing_parse_fastq_record(text, ls, n_lines, hdr[k], file, replace_n, r);
    const uint64_t pos = ing_fastq_pos(k, file, phase);     // mate 1 / mate 2 alternate (fastqs_to_sam.cpp:49)
    recs[pos] = r;
    ing_report(err, pos, r.err);
  }
}
int launch_ing_parse_fastq(const uint8_t *text, const uint64_t *ls, uint64_t n_lines, const uint64_t *hdr, uint64_t n_take, int file,
                           int phase, int replace_n, LineRec *recs, unsigned long long *err, cudaStream_t st) {
  if (!n_take) return 0;
  const uint64_t want = (n_take + 127) / 128;
  const unsigned grid = (unsigned)(want < 148ull * 64 ? want : 148ull * 64);
  ing_launch(k_ing_parse_fastq, grid, 128, st, text, ls, n_lines, hdr, n_take, file, phase, replace_n, recs, err);
  return 1;
}

// ---- ordered compaction ----------------------------------------------------------------------------
struct RecIn {
  const LineRec *recs;
  __device__ __forceinline__ Ing4 operator()(uint64_t i) const { return ing4_of(recs[i]); }
};
struct PreOut {
  Ing4 *pre;
  __device__ __forceinline__ void operator()(uint64_t i, const Ing4 &ex, const Ing4 &) const { pre[i] = ex; }
  __device__ __forceinline__ void total(uint64_t n, const Ing4 &tot) const { pre[n] = tot; }
};
int launch_ing_scan_recs(const LineRec *recs, uint64_t m, Ing4 *blk4, Ing4 *pre, cudaStream_t st) {
  return scan3<Ing4>(RecIn{recs}, m, blk4, OpAdd4(), Ing4{0, 0, 0, 0}, PreOut{pre}, st);
}

// Totals of the batch, the pairing rule at a chunk edge and the bytes consumed, published to mapped host memory.
// An odd number of reads in a non-final chunk would shift the reader's pairing by arrival parity
// (query.cpp:629-637) for everything after it: the last read is left for the next chunk.
__global__ void k_ing_publish(IngPublish p) {
  if (threadIdx.x || blockIdx.x) return;
  uint64_t m = p.m;
  Ing4 tot = p.pre[m];
  if (!p.final && (tot.reads & 1ull)) {
    uint64_t lo = 0, hi = m;                                // smallest q with pre[q + 1].reads == tot.reads: the last read
    while (lo < hi) { const uint64_t mid = lo + (hi - lo) / 2; if (p.pre[mid + 1].reads >= tot.reads) hi = mid; else lo = mid + 1; }
    m = lo;
    tot = p.pre[m];
  }
  uint64_t consumed[2] = {p.n_bytes[0], p.n_bytes[1]};
  if (p.fastq) {
    for (int f = 0; f < 2; ++f) {
      const uint64_t take = ing_fastq_taken(m, f, p.phase);
      if (take < p.n_rec[f]) consumed[f] = p.ls[f][p.hdr[f][take]];
    }
  } else {
    if (m < p.n_rec[0]) consumed[0] = p.ls[0][m];
    consumed[1] = 0;
  }
  volatile uint64_t *h = p.host;
  h[0] = tot.reads; h[1] = tot.name; h[2] = tot.seq; h[3] = tot.opt; h[4] = m; h[5] = consumed[0]; h[6] = consumed[1];
  h[7] = (uint64_t)*p.err;
  h[12] = (uint64_t)((p.phase ^ (int)(m & 1ull)) & 1);      // FASTQ: which mate file the next chunk starts with
  __threadfence_system();
}
int launch_ing_publish(const IngPublish &p, cudaStream_t st) {
  ing_launch(k_ing_publish, 1, 32, st, p);
  return 1;
}

// One run of bytes (SEQ or QUAL) copied by the 8 lanes of a read's group: the destination's aligned 16-byte chunks
// are assembled from five aligned source words by funnel shifts, the < 16 bytes either side go out as byte stores.
// fix = 1 replaces 'N' by 'Z' (fastqs_to_sam.cpp:69).  May read up to 7 bytes beyond the run (buffer slack).
__device__ __forceinline__ uint32_t ing_n2z_word(uint32_t w) {
  const uint32_t x = w ^ 0x4e4e4e4eu;                                   // 'N'
  const uint32_t t = ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x;
  return w ^ (((~t & 0x80808080u) >> 7) * 0x14u);                       // 'N' ^ 'Z' == 0x14
}
__device__ __forceinline__ void ing_copy_run8(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint32_t len, int l, bool fix) {
  const uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
  const uint32_t hl = head < len ? head : len;
  const uint32_t nb = (len - hl) >> 4, tail = (len - hl) & 15u, t0 = hl + 16u * nb;
  for (uint32_t k = l; k < hl; k += 8) { uint8_t ch = src[k]; if (fix && ch == 'N') ch = 'Z'; dst[k] = ch; }
  for (uint32_t k = l; k < tail; k += 8) { uint8_t ch = src[t0 + k]; if (fix && ch == 'N') ch = 'Z'; dst[t0 + k] = ch; }
  for (uint32_t k = l; k < nb; k += 8) {
    const uint32_t s0 = hl + 16u * k;
    const uintptr_t a = (uintptr_t)(src + s0);
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8u;
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
    uint4 v;
    v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
    v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
    if (fix) { v.x = ing_n2z_word(v.x); v.y = ing_n2z_word(v.y); v.z = ing_n2z_word(v.z); v.w = ing_n2z_word(v.w); }
    *reinterpret_cast<uint4 *>(dst + s0) = v;
  }
}

// 8 lanes per read, 4 reads per warp in flight; no warp-level collectives (the groups diverge freely).
__global__ void __launch_bounds__(256) k_ing_copy(IngCopy c) {
  const int l = threadIdx.x & 7;
  const uint64_t group0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const uint64_t n_groups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
  if (group0 == 0 && l == 0) {                              // end sentinels of the offset arrays
    const Ing4 tot = c.pre[c.m];
    c.name_off[tot.reads] = (int64_t)tot.name; c.seq_off[tot.reads] = (int64_t)tot.seq;
    if (c.opt_off) c.opt_off[tot.reads] = (int64_t)tot.opt;
  }
  for (uint64_t i = group0; i < c.m; i += n_groups) {
    const LineRec r = c.recs[i];
    if (!(r.bits & ING_EMIT)) continue;
    const Ing4 p = c.pre[i];
    const uint8_t *__restrict__ t = c.text[r.src];
    if (l == 0) {
      c.name_off[p.reads] = (int64_t)p.name; c.seq_off[p.reads] = (int64_t)p.seq; c.read_flag[p.reads] = r.read_flag;
      if (c.opt_off) c.opt_off[p.reads] = (int64_t)p.opt;
    }
    for (uint32_t k = l; k < r.name_len; k += 8) c.names[p.name + k] = t[r.name_pos + k];
    ing_copy_run8(c.seq + p.seq, t + r.seq_pos, r.seq_len, l, (r.bits & ING_N2Z) != 0);
    ing_copy_run8(c.qual + p.seq, t + r.qual_pos, r.seq_len, l, false);
    if (r.bits & ING_OPT_SAM) {
      const uint64_t b = r.opt_pos - 1, e = b + r.opt_src_len;
      if (r.opt_len == r.opt_src_len) {
        // no whitespace run to collapse: every source byte yields one output byte (separators become tabs)
        for (uint32_t k = l; k < r.opt_src_len; k += 8) { const uint8_t ch = t[b + k]; c.opt[p.opt + k] = ing_space(ch) ? (uint8_t)'\t' : ch; }
      } else if (l == 0) {
        uint64_t o = p.opt;
        for (uint64_t k = b; k < e; ++k) { const int ch = ing_opt_char(t, k, e); if (ch >= 0) c.opt[o++] = (uint8_t)ch; }
      }
    } else if (r.bits & ING_OPT_XO) {
      const uint64_t lit = 0x3a5a3a4f5809ull;                // "\tXO:Z:" little-endian (fastqs_to_sam.cpp:88-91 + add_optional's tab)
      if (l < 6) c.opt[p.opt + l] = (uint8_t)(lit >> (8 * l));
      for (uint32_t k = l; k < r.opt_src_len; k += 8) c.opt[p.opt + 6 + k] = t[r.opt_pos + k];
    }
  }
}
int launch_ing_copy(const IngCopy &c, cudaStream_t st) {
  const uint64_t want = (c.m + 31) / 32;                     // 32 groups of 8 lanes per block
  const unsigned grid = (unsigned)(want < 148ull * 16 ? (want ? want : 1) : 148ull * 16);
  ing_launch(k_ing_copy, grid, 256, st, c);
  return 1;
}

}  // namespace smash
