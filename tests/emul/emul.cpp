// tests/emul/emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host-side, lane-by-lane emulation of the kernels in smash_paper_b200/csrc/kernels.cu built from
// the very same __host__ __device__ building blocks (core.cuh, records.cuh).  It exists so the CPU
// test-suite (no GPU in the dev container) can exercise the kernels' logic against the oracle.
// It is never linked into libsmash_b200.so and nothing in the product calls it.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../smash_paper_b200/csrc/core.cuh"
#include "../../smash_paper_b200/csrc/records.cuh"

using namespace smash;

struct EmulIndex {
  DevIndex d;
  std::vector<uint8_t> text_padded, uniq;
  std::vector<uint64_t> seed;
  std::vector<LcpItem> lcpm;
  std::vector<int> descr_off;
  std::string descr;
  std::vector<uint32_t> off32;
  std::vector<uint8_t> mapbody;
  std::vector<uint32_t> ext;
  std::vector<uint64_t> descr8;
};

static uint64_t kmers_le_suffix(const DevIndex &ix, uint64_t i, int k) {   // mirrors kernels.cu
  const uint64_t s = sa_at(ix, i);
  uint64_t code = 0;
  for (int j = 0; j < k; ++j) {
    const uint64_t p = s + (uint64_t)j;
    const uint8_t ch = p < ix.N ? ix.text[p] : 0;
    const int b = base_code(ch);
    if (b <= 3) { code = (code << 2) | (uint64_t)b; continue; }
    const uint64_t below = ch < 'a' ? 0 : ch < 'c' ? 1 : ch < 'g' ? 2 : ch < 't' ? 3 : 4;
    return ((code << 2) + below) << (2 * (k - j - 1));
  }
  return code + 1;
}

extern "C" {

void *emul_index_create(const uint8_t *text, uint64_t N, const void *sa, const void *isa, int w,
                        const uint8_t *lcp, const uint8_t *lcp_m_raw, uint64_t n_m, uint64_t n_descr,
                        const uint64_t *startpos, const uint64_t *sizes, const char *const *descr,
                        int rcref, int seed_k, const uint8_t *mapbody, uint64_t map_bytes) {
  EmulIndex *e = new EmulIndex();
  e->text_padded.assign(N + 2 * TEXT_PAD, 0);
  memcpy(e->text_padded.data() + TEXT_PAD, text, N);
  DevIndex &d = e->d;
  memset(&d, 0, sizeof d);
  d.text = e->text_padded.data() + TEXT_PAD; d.N = N; d.sa = sa; d.isa = isa; d.w = w; d.lcp = lcp;
  e->lcpm.resize(n_m + 1);
  for (uint64_t i = 0; i < n_m; ++i) {
    memcpy(&e->lcpm[i].idx, lcp_m_raw + 16 * i, 8);
    if (w == 4) { uint32_t v; memcpy(&v, lcp_m_raw + 16 * i + 8, 4); e->lcpm[i].val = v; }
    else memcpy(&e->lcpm[i].val, lcp_m_raw + 16 * i + 8, 8);
  }
  d.lcp_m = e->lcpm.data(); d.n_m = n_m;
  d.startpos = startpos; d.sizes = sizes; d.n_descr = (int)n_descr; d.rcref = rcref;
  for (uint64_t i = 0; i < n_descr; ++i) { e->descr_off.push_back((int)e->descr.size()); e->descr += descr[i]; }
  e->descr_off.push_back((int)e->descr.size());
  d.descr = e->descr.data(); d.descr_off = e->descr_off.data();
  for (uint64_t i = 0; i < n_descr; ++i) { uint64_t v = 0; for (size_t j = 0; j < strlen(descr[i]) && j < 8; ++j) v |= (uint64_t)(uint8_t)descr[i][j] << (8 * j); e->descr8.push_back(v); }
  d.descr8 = e->descr8.data();
  d.logN = (uint64_t)ceil(log((double)N) / log(2.0));
  for (uint64_t i = 0; i < N; ++i) d.alpha[text[i] >> 5] |= 1u << (text[i] & 31);
  e->uniq.assign(N, 0);
  for (uint64_t i = 0; i < N; ++i) {
    int a = lcp[i], b = i + 1 < N ? lcp[i + 1] : 0, m = a > b ? a : b;
    e->uniq[sa_at(d, i)] = (uint8_t)(m >= 254 ? 255 : m + 1);
  }
  d.uniq = e->uniq.data();
  int k = seed_k;
  if (k <= 0) k = (int)ceil(log((double)N) / log(4.0)) + 1;
  if (k > 16) k = 16;
  if (k < 4) k = 4;
  e->seed.assign((1ull << (2 * k)) + 1, 0);
  for (uint64_t i = 0; i <= N; ++i) {
    const uint64_t prev = i ? kmers_le_suffix(d, i - 1, k) : 0;
    const uint64_t cur = i < N ? kmers_le_suffix(d, i, k) : (1ull << (2 * k)) + 1;
    for (uint64_t x = prev; x < cur; ++x) e->seed[x] = i;
  }
  d.seed = e->seed.data(); d.seed_k = k; d.seed_w = 8;
  uint32_t acc = 0;
  for (uint64_t i = 0; i < n_descr; i += rcref ? 2 : 1) { e->off32.push_back(acc); acc += (uint32_t)sizes[i]; }
  e->off32.push_back(acc);
  d.chrom_off32 = e->off32.data();
  {
    // 8+6 character pre-filter codes (core.cuh ext_entry, what k_ext_build stores)
    e->ext.assign(N, 0);
    for (uint64_t i = 0; i < N; ++i) e->ext[i] = ext_entry(text, N, sa_at(d, i), d.seed_k);
  }
  d.ext = e->ext.data();
  if (mapbody && map_bytes) { e->mapbody.assign(mapbody, mapbody + map_bytes); d.mapbody = e->mapbody.data(); d.map_bytes = map_bytes; }
  return e;
}
void emul_index_destroy(void *p) { delete (EmulIndex *)p; }

// One "warp": stage the read, run anchors lane by lane (or the exact path), rank-sort the stage.
static int emul_search(const DevIndex &ix, const uint8_t *seq, int q, uint32_t min_len, int nuc_only,
                       int force_exact, std::vector<uint8_t> &pbuf, std::vector<Match> &out) {
  pbuf.assign(P_FRONT + q + P_BACK + 8, 0);
  uint8_t *P = pbuf.data() + P_FRONT;
  bool odd = false;
  for (int j = 0; j < q; ++j) { uint8_t c = query_char(seq[j], nuc_only); P[j] = c; if (base_code(c) > 3 && in_alpha(ix, c)) odd = true; }
  for (int j = 0; j < P_FRONT; ++j) pbuf[j] = 0xFE;
  for (int j = 0; j < P_BACK; ++j) P[q + j] = 0xFF;
  const uint32_t L = min_len < 2 ? 2 : min_len;
  const int k = ix.seed_k < (int)L ? ix.seed_k : (int)L;
  const int s = (int)L - k + 1;
  const double expect = (double)ix.N / pow(4.0, (double)k);
  const bool fast_ok = expect <= 16.0 && !force_exact;
  std::vector<Match> stage;
  if (q >= (int)L) {
    if (!odd && fast_ok) {
      const int n_anchor = (q - (int)L + s - 1) / s + 1;
      std::vector<uint32_t> inv(q / 32 + 3, 0);
      for (int j = 0; j < q; ++j) if (base_code(P[j]) > 3) inv[j >> 5] |= 1u << (j & 31);
      for (int a = 0; a < n_anchor; ++a) {
        const int x = a * s;
        if (kmer_invalid(inv.data(), x, k)) continue;
        uint64_t lo, hi;
        anchor_bucket(ix, P, x, k, &lo, &hi);
        if (hi - lo > (uint64_t)BIG_BUCKET) {
          const int p_lo = x - s + 1 > 0 ? x - s + 1 : 0;
          for (int p = p_lo; p <= x; ++p) { Match m; if (exact_start(ix, P, q, p, L, &m)) stage.push_back(m); }
          continue;
        }
        for (uint64_t i = lo; i < hi; ++i) {
          if (k == ix.seed_k && ix.ext && !ext_may_reach(ix.ext[i], read_ext_codes(P, x, k), k, L)) continue;
          Match m; int pl = 0;
          const int r = candidate_check(ix, P, q, x, s, k, L, sa_at(ix, i), &m, &pl);
          if (r > 0) stage.push_back(m);
          else if (r < 0 && exact_start(ix, P, q, pl, L, &m)) stage.push_back(m);
        }
      }
    } else {
      for (int p = 0; p + (int)L <= q; ++p) { Match m; if (exact_start(ix, P, q, p, L, &m)) stage.push_back(m); }
    }
  }
  out.assign(stage.size(), Match{});
  for (size_t e = 0; e < stage.size(); ++e) {
    size_t rank = 0;
    for (size_t f = 0; f < stage.size(); ++f) rank += stage[f].qpos < stage[e].qpos;
    out[rank] = stage[e];
  }
  return (int)out.size();
}

struct ReadOut { std::vector<Item> items; std::vector<Rec> recs; ReadSum sum; std::vector<uint8_t> pbuf; };

// Returns SAM bytes needed; fills out if cap suffices. matches_out: triples (ref,q,len), match_off n+1.
uint64_t emul_map_batch(void *index, uint64_t n_reads, const uint8_t *names, const int64_t *name_off,
                        const uint8_t *seq, const uint8_t *qual, const int64_t *seq_off,
                        const uint8_t *opt, const int64_t *opt_off, const uint16_t *read_flag,
                        uint32_t min_len, int nomap, int nuc_only, int tag_map, int force_exact,
                        char *out, uint64_t cap, int64_t *match_off, uint64_t *match_triples, uint64_t match_cap,
                        uint32_t *maperr) {
  EmulIndex *e = (EmulIndex *)index;
  const DevIndex &ix = e->d;
  std::vector<ReadOut> ro(n_reads);
  uint64_t mtot = 0;
  if (maperr) *maperr = 0;
  for (uint64_t r = 0; r < n_reads; ++r) {
    const int q = (int)(seq_off[r + 1] - seq_off[r]);
    std::vector<Match> m;
    emul_search(ix, seq + seq_off[r], q, min_len, nuc_only, force_exact, ro[r].pbuf, m);
    if (match_off) {
      match_off[r] = (int64_t)mtot;
      for (auto &x : m) { if (mtot < match_cap) { match_triples[3 * mtot] = x.ref; match_triples[3 * mtot + 1] = x.qpos; match_triples[3 * mtot + 2] = x.len; } ++mtot; }
    }
    const int n = (int)m.size();
    std::vector<Aln> aln(n + 1); std::vector<uint16_t> ord(n + 1);
    ro[r].items.resize(n + 1); ro[r].recs.resize(n + 1);
    build_records(ix, m.data(), n, q, nomap, aln.data(), ord.data(), ro[r].items.data(), ro[r].recs.data(), &ro[r].sum);
    const uint8_t *P = ro[r].pbuf.data() + P_FRONT;
    if (!ro[r].sum.unmapped)
      for (int k = 0; k < ro[r].sum.n_rec; ++k) {
        int cnt = 0;
        for (int j0 = 0; j0 < q; j0 += 8) cnt += xe_word(ix, P, q, ro[r].recs[k].rcpos, j0);
        ro[r].recs[k].xe = (uint16_t)cnt;
        if (ix.mapbody) {
          bool ok = true;
          for (int u = 0; u < ro[r].recs[k].item_cnt; ++u) {
            Item it = ro[r].items[ro[r].recs[k].item_begin + u]; int L, R;
            ok = map_lr(ix, ro[r].recs[k].si >> 1, ro[r].recs[k].pos, it.prefix, it.len, &L, &R) && ok;
            ro[r].items[ro[r].recs[k].item_begin + u].L = (uint8_t)L; ro[r].items[ro[r].recs[k].item_begin + u].R = (uint8_t)R;
            if (u == 0) { ro[r].recs[k].L0 = (uint8_t)L; ro[r].recs[k].R0 = (uint8_t)R; }
          }
          if (!ok && maperr) ++*maperr;
        }
      }
  }
  if (match_off) match_off[n_reads] = (int64_t)mtot;
  std::string sam;
  for (uint64_t r = 0; r < n_reads; ++r) {
    const ReadSum &me = ro[r].sum;
    if (!me.n_rec) continue;
    uint16_t flag; MateView mv;
    const uint16_t mine = (uint16_t)(read_flag[r] | (me.unmapped ? 4 : 0));
    const uint64_t other = r ^ 1ull;
    if (other < n_reads) {
      const ReadSum &ot = ro[other].sum;
      mate_view(mine, me, (uint16_t)(read_flag[other] | (ot.unmapped ? 4 : 0)), &ot, (r & 1) == 0, &mv, &flag);
    } else mate_view(mine, me, 0, nullptr, true, &mv, &flag);
    const int q = (int)(seq_off[r + 1] - seq_off[r]);
    const char *name = (const char *)names + name_off[r];
    const int name_len = (int)(name_off[r + 1] - name_off[r]);
    for (int k = 0; k < me.n_rec; ++k) {
      std::vector<char> buf(name_len + 4096 + 64 * ro[r].items.size());
      BufSink bs{buf.data()};
      put_head(bs, ix, name, name_len, flag, me.unmapped, ro[r].recs.data(), k, me.n_rec, ro[r].items.data(), mv);
      CountSink cs; put_head(cs, ix, name, name_len, flag, me.unmapped, ro[r].recs.data(), k, me.n_rec, ro[r].items.data(), mv);
      if (cs.n != bs.n) abort();
      sam.append(buf.data(), bs.n);
      const uint8_t *sq = seq + seq_off[r], *ql = qual + seq_off[r];
      if (ro[r].recs[k].rc && !me.unmapped) {
        for (int j = 0; j < q; ++j) sam.push_back((char)comp_char(sq[q - 1 - j]));
        sam.push_back('\t');
        for (int j = 0; j < q; ++j) sam.push_back((char)ql[q - 1 - j]);
      } else { sam.append((const char *)sq, q); sam.push_back('\t'); sam.append((const char *)ql, q); }
      BufSink ts{buf.data()};
      put_tags(ts, ix, me.unmapped, ro[r].recs.data(), k, me.n_rec, ro[r].items.data());
      sam.append(buf.data(), ts.n);
      if (opt) sam.append((const char *)opt + opt_off[r], (size_t)(opt_off[r + 1] - opt_off[r]));
      if (tag_map && !me.unmapped) { BufSink ls{buf.data()}; put_lr_tags(ls, ix, ro[r].recs[k], ro[r].items.data()); sam.append(buf.data(), ls.n); }
      sam.push_back('\n');
    }
  }
  if (out && sam.size() <= cap) memcpy(out, sam.data(), sam.size());
  return sam.size();
}

// WordSink (aligned 8-byte stores at arbitrary byte alignment) against the plain byte sink.
int emul_wordsink_selftest() {
  alignas(8) char a[256], b[256];
  for (int off = 0; off < 9; ++off)
    for (int len = 0; len < 40; ++len) {
      memset(a, '.', sizeof a); memset(b, '.', sizeof b);
      BufSink bs{b + 16 + off};
      WordSink ws(a + 16 + off);
      for (int i = 0; i < len; ++i) { bs.ch((char)('A' + i % 26)); ws.ch((char)('A' + i % 26)); }
      put_u64(bs, 1234567890123ull + (uint64_t)len); put_u64(ws, 1234567890123ull + (uint64_t)len);
      put_lit(bs, "\tXM:i:0\tNH:i:0"); put_lit(ws, "\tXM:i:0\tNH:i:0"); put_u64(bs, (uint64_t)len * 7919u); put_u64(ws, (uint64_t)len * 7919u);
      put_u64(bs, 100000000ull * (uint64_t)off + 42); put_u64(ws, 100000000ull * (uint64_t)off + 42);
      put_i64(bs, -(int64_t)off); put_i64(ws, -(int64_t)off);
      ws.finish();
      if (ws.n != bs.n || memcmp(a, b, sizeof a) != 0) return 1 + off * 100 + len;
    }
  return 0;
}

}  // extern "C"
