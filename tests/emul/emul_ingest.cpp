// tests/emul/emul_ingest.cpp -- TEST INFRASTRUCTURE ONLY.
// Host-side emulation of the device-side input stage (smash_paper_b200/csrc/ingest.cu) built from the same
// __host__ __device__ building blocks (ingest.cuh), stage by stage in the order the kernels run:
// chunk line-start counts -> line starts -> [FASTQ: transition-function scan, header compaction] ->
// LineRec per line/record -> Ing4 scan -> publish rule -> copy.  It lets the CPU test-suite check the
// kernels' logic against oracle/ingest.py where no GPU exists.  Never linked into libsmash_b200.so.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../smash_paper_b200/csrc/ingest.cuh"

using namespace smash;

namespace {

struct Lines { std::vector<uint64_t> ls; uint64_t n_lines = 0; };

static inline uint64_t word_at(const uint8_t *p) { uint64_t w; memcpy(&w, p, 8); return w; }

// k_ing_scan_tiles/top/apply with ChunkIn / LineStartOut.  text: padded copy (64 bytes of slack, like ing_raw)
Lines line_starts(const uint8_t *text, uint64_t n) {
  Lines L;
  if (!n) return L;
  const uint64_t n_chunks = (n + 15) / 16;
  std::vector<uint64_t> cnt(n_chunks), pre(n_chunks + 1, 0);
  for (uint64_t c = 0; c < n_chunks; ++c) cnt[c] = ing_chunk_starts(word_at(text + 16 * c), word_at(text + 16 * c + 8), 16 * c, n);
  for (uint64_t c = 0; c < n_chunks; ++c) pre[c + 1] = pre[c] + cnt[c];
  L.n_lines = pre[n_chunks];
  L.ls.assign(L.n_lines + 1, 0);
  for (uint64_t c = 0; c < n_chunks; ++c) {
    if (!cnt[c]) continue;
    const uint64_t end = ing_chunk_place(word_at(text + 16 * c), word_at(text + 16 * c + 8), 16 * c, n, L.ls.data(), pre[c]);
    if (end != pre[c + 1]) abort();
  }
  L.ls[L.n_lines] = ing_sentinel(text, n);
  return L;
}

// FsmIn/FsmOut + FlagIn/HdrOut
std::vector<uint64_t> fastq_headers(const uint8_t *text, const Lines &L) {
  std::vector<uint64_t> hdr;
  uint32_t before = FQ_IDENT;
  for (uint64_t j = 0; j < L.n_lines; ++j) {
    const uint32_t fn = fq_line_fn(ing_first_char(text, L.ls[j], ing_line_end(L.ls.data(), j)));
    if (fq_apply(before, FQ_H) == FQ_H && fq_apply(fn, FQ_H) != FQ_H) hdr.push_back(j);
    before = fq_compose(before, fn);
  }
  return hdr;
}

}  // namespace

extern "C" {

// SWAR primitives against their byte-wise definitions: every byte value at every position, random words, and
// ing_load8 / ing_token_end at every alignment.  Returns the number of disagreements.
int emul_ingest_swar_check(void) {
  int bad = 0;
  uint64_t x = 0x9E3779B97F4A7C15ull;
  for (int it = 0; it < 200000; ++it) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    uint64_t w = x;
    if (it < 256 * 8) { const int pos = it & 7; w = (w & ~(0xffull << (8 * pos))) | ((uint64_t)(it >> 3) << (8 * pos)); }
    if (it & 1) w &= 0x3f3f3f3f3f3f3f3full;                   // more low bytes: whitespace and newlines show up
    uint64_t ws = 0, nl = 0;
    for (int b = 0; b < 8; ++b) {
      const uint8_t c = (uint8_t)(w >> (8 * b));
      if (ing_space(c)) ws |= 0x80ull << (8 * b);
      if (c == 10) nl |= 0x80ull << (8 * b);
    }
    if (ing_ws_mask(w) != ws) ++bad;
    if (ing_eq_mask(w, 10) != nl) ++bad;
  }
  alignas(16) uint8_t buf[96];
  for (int i = 0; i < 96; ++i) buf[i] = (uint8_t)(33 + (i * 7) % 90);
  for (int off = 0; off < 24; ++off) {
    uint64_t w; memcpy(&w, buf + off, 8);
    if (ing_load8(buf, off) != w) ++bad;
    for (int len = 0; len < 40; ++len) {
      uint8_t t[96]; memcpy(t, buf, 96);
      t[off + len] = (len % 3 == 0) ? ' ' : (len % 3 == 1 ? '\t' : '\r');
      if (ing_token_end(t, off, 80) != (uint64_t)(off + len)) ++bad;
      if (ing_token_end(t, off, off + len) != (uint64_t)(off + len)) ++bad;      // token runs to e
    }
  }
  return bad;
}

// kind 0: SAM text (text1 ignored), 1: FASTQ pair.  Outputs are caller buffers sized generously
// (names/seq/qual/opt: n0 + n1 bytes; offsets / read_flag: line count + 1).  Returns 0, or the IngErr code
// with *err_index = record index.  out_info: n_reads, name_bytes, seq_bytes, opt_bytes, consumed0, consumed1, next phase.
int emul_ingest(int kind, int final, int replace_n, int phase, const uint8_t *text0, uint64_t n0, const uint8_t *text1, uint64_t n1,
                uint8_t *names, int64_t *name_off, uint8_t *seq, uint8_t *qual, int64_t *seq_off, uint8_t *opt, int64_t *opt_off,
                uint16_t *read_flag, uint64_t *out_info, uint64_t *err_index) {
  phase = kind ? (phase & 1) : 0;
  const uint8_t *given[2] = {text0, text1};
  const uint64_t full[2] = {n0, kind ? n1 : 0};
  uint64_t nb[2] = {n0, kind ? n1 : 0};
  if (!final) for (int f = 0; f < 2; ++f) while (nb[f] && given[f][nb[f] - 1] != '\n') --nb[f];
  std::vector<uint8_t> padded[2];                            // the device buffers: garbage past the text, 64 bytes of slack
  const uint8_t *text[2];
  for (int f = 0; f < 2; ++f) { padded[f].assign(nb[f] + 64, 0xAB); if (nb[f]) memcpy(padded[f].data(), given[f], nb[f]); text[f] = padded[f].data(); }
  Lines L[2];
  std::vector<uint64_t> hdr[2];
  uint64_t n_rec[2] = {0, 0}, n_take[2] = {0, 0};
  for (int f = 0; f < (kind ? 2 : 1); ++f) L[f] = line_starts(text[f], nb[f]);
  std::vector<LineRec> recs;
  uint64_t m = 0;
  if (kind) {
    for (int f = 0; f < 2; ++f) { hdr[f] = fastq_headers(text[f], L[f]); n_rec[f] = hdr[f].size(); }
    ing_fastq_take(n_rec[phase], n_rec[phase ^ 1], final, &n_take[phase], &n_take[phase ^ 1]);
    m = n_take[0] + n_take[1];
    recs.resize(m + 1);
    for (int f = 0; f < 2; ++f)
      for (uint64_t k = 0; k < n_take[f]; ++k)
        ing_parse_fastq_record(text[f], L[f].ls.data(), L[f].n_lines, hdr[f][k], f, replace_n, recs[ing_fastq_pos(k, f, phase)]);
  } else {
    n_rec[0] = L[0].n_lines;
    m = L[0].n_lines;
    recs.resize(m + 1);
    for (uint64_t j = 0; j < m; ++j) ing_parse_sam_line(text[0], L[0].ls[j], ing_line_end(L[0].ls.data(), j), recs[j]);
  }
  uint64_t err = ~0ull;
  for (uint64_t i = 0; i < m; ++i) if (recs[i].err) { const uint64_t v = (i << 8) | recs[i].err; if (v < err) err = v; }
  // Ing4 scan
  std::vector<Ing4> pre(m + 1);
  Ing4 acc{0, 0, 0, 0};
  for (uint64_t i = 0; i < m; ++i) { pre[i] = acc; acc = ing4_add(acc, ing4_of(recs[i])); }
  pre[m] = acc;
  // k_ing_publish
  uint64_t mu = m;
  Ing4 tot = pre[m];
  if (m && !final && (tot.reads & 1ull)) {
    uint64_t lo = 0, hi = m;
    while (lo < hi) { const uint64_t mid = lo + (hi - lo) / 2; if (pre[mid + 1].reads >= tot.reads) hi = mid; else lo = mid + 1; }
    mu = lo; tot = pre[mu];
  }
  uint64_t consumed[2] = {nb[0], nb[1]};
  if (m) {
    if (kind) {
      for (int f = 0; f < 2; ++f) { const uint64_t take = ing_fastq_taken(mu, f, phase); if (take < n_rec[f]) consumed[f] = L[f].ls[hdr[f][take]]; }
    } else {
      if (mu < n_rec[0]) consumed[0] = L[0].ls[mu];
      consumed[1] = 0;
    }
  } else {
    consumed[0] = final ? full[0] : 0; consumed[1] = final ? full[1] : 0;
  }
  if (final) { consumed[0] = full[0]; consumed[1] = full[1]; }
  if (err != ~0ull) { *err_index = err >> 8; return (int)(err & 0xff); }
  out_info[0] = tot.reads; out_info[1] = tot.name; out_info[2] = tot.seq; out_info[3] = tot.opt; out_info[4] = consumed[0]; out_info[5] = consumed[1]; out_info[6] = m ? (uint64_t)((phase ^ (int)(mu & 1)) & 1) : (uint64_t)phase;
  // k_ing_copy
  name_off[tot.reads] = (int64_t)tot.name; seq_off[tot.reads] = (int64_t)tot.seq; opt_off[tot.reads] = (int64_t)tot.opt;
  for (uint64_t i = 0; i < mu; ++i) {
    const LineRec &r = recs[i];
    if (!(r.bits & ING_EMIT)) continue;
    const Ing4 p = pre[i];
    const uint8_t *t = text[r.src];
    name_off[p.reads] = (int64_t)p.name; seq_off[p.reads] = (int64_t)p.seq; opt_off[p.reads] = (int64_t)p.opt; read_flag[p.reads] = r.read_flag;
    for (uint32_t k = 0; k < r.name_len; ++k) names[p.name + k] = t[r.name_pos + k];
    for (uint32_t k = 0; k < r.seq_len; ++k) {
      uint8_t ch = t[r.seq_pos + k];
      if ((r.bits & ING_N2Z) && ch == 'N') ch = 'Z';
      seq[p.seq + k] = ch;
      qual[p.seq + k] = t[r.qual_pos + k];
    }
    if (r.bits & ING_OPT_SAM) {
      const uint64_t b = r.opt_pos - 1, e = b + r.opt_src_len;
      uint64_t o = p.opt;
      for (uint64_t k = b; k < e; ++k) { const int ch = ing_opt_char(t, k, e); if (ch >= 0) opt[o++] = (uint8_t)ch; }
      if (o != p.opt + r.opt_len) return 100;                // internal: the size pass and the copy disagree
    } else if (r.bits & ING_OPT_XO) {
      const uint64_t lit = 0x3a5a3a4f5809ull;
      for (int k = 0; k < 6; ++k) opt[p.opt + k] = (uint8_t)(lit >> (8 * k));
      for (uint32_t k = 0; k < r.opt_src_len; ++k) opt[p.opt + 6 + k] = t[r.opt_pos + k];
    }
  }
  return 0;
}

}  // extern "C"
