// tests/emul/emul_ingest_kernels.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles smash_paper_b200/csrc/ingest.cu ITSELF with g++ against tests/emul/cuda_shim and EXECUTES its kernels on
// the host (blocks of OS threads, real barriers, real shuffles through per-warp slots), driven by the same
// sequence of launches as api.cu's ingest_text.  Where emul_ingest.cpp checks the parsing logic, this checks the
// kernels' indexing, scans, shuffles and launch geometry without a GPU.  Same C interface as emul_ingest().
#include "cuda_shim/cuda_runtime.h"

#include "../../smash_paper_b200/csrc/ingest.cu"

#include <vector>

using namespace smash;

extern "C" int emul_ingest_kernels(int kind, int final, int replace_n, int phase, const uint8_t *text0, uint64_t n0, const uint8_t *text1, uint64_t n1,
                                   uint8_t *names, int64_t *name_off, uint8_t *seq, uint8_t *qual, int64_t *seq_off, uint8_t *opt,
                                   int64_t *opt_off, uint16_t *read_flag, uint64_t *out_info, uint64_t *err_index) {
  const bool fastq = kind != 0;
  phase = fastq ? (phase & 1) : 0;
  const int n_text = fastq ? 2 : 1;
  const uint8_t *text[2] = {text0, text1};
  const uint64_t full[2] = {n0, fastq ? n1 : 0};
  uint64_t nb[2] = {n0, fastq ? n1 : 0};
  if (!final) for (int f = 0; f < 2; ++f) while (nb[f] && text[f][nb[f] - 1] != '\n') --nb[f];
  cudaStream_t st = nullptr;
  // "device" buffers
  std::vector<uint8_t> raw[2];
  std::vector<uint64_t> blk_lines[2], ls[2], hdr[2];
  std::vector<uint32_t> blk32;
  std::vector<uint8_t> hdr_flag;
  unsigned long long scal[3] = {~0ull, 0, 0};
  uint64_t n_lines[2] = {0, 0};
  for (int f = 0; f < n_text; ++f) {
    if (!nb[f]) continue;
    raw[f].assign(nb[f] + 64, 0xAB);
    memcpy(raw[f].data(), text[f], nb[f]);
    const uint64_t tiles = ing_tiles((nb[f] + 15) / 16);
    blk_lines[f].assign(tiles + 2, 0);
    launch_ing_count_lines(raw[f].data(), nb[f], blk_lines[f].data(), st);
    n_lines[f] = blk_lines[f][tiles];
  }
  for (int f = 0; f < n_text; ++f) {
    if (!n_lines[f]) continue;
    ls[f].assign(n_lines[f] + 2, 0);
    launch_ing_line_starts(raw[f].data(), nb[f], blk_lines[f].data(), ls[f].data(), st);
  }
  uint64_t m = 0, n_rec[2] = {n_lines[0], 0}, n_take[2] = {0, 0};
  if (fastq) {
    for (int f = 0; f < 2; ++f) {
      if (!n_lines[f]) continue;
      hdr_flag.assign(n_lines[f] + 1, 0);
      blk32.assign(ing_tiles(n_lines[f]) + 2, 0);
      hdr[f].assign(n_lines[f] + ing_tiles(n_lines[f]) + 4, 0);
      launch_ing_fastq_headers(raw[f].data(), ls[f].data(), n_lines[f], blk32.data(), hdr[f].data() + n_lines[f] + 1, hdr_flag.data(),
                               hdr[f].data(), (uint64_t *)&scal[1 + f], st);
    }
    n_rec[0] = scal[1]; n_rec[1] = scal[2];
    ing_fastq_take(n_rec[phase], n_rec[phase ^ 1], final, &n_take[phase], &n_take[phase ^ 1]);
    m = n_take[0] + n_take[1];
  } else {
    m = n_lines[0];
  }
  uint64_t host[16] = {0, 0, 0, 0, 0, 0, 0, ~0ull, 0, 0, 0, 0, (uint64_t)phase};
  uint64_t consumed[2] = {0, 0};
  std::vector<LineRec> recs(m + 1);
  std::vector<Ing4> pre(m + 2), blk4(ing_tiles(m) + 2);
  if (m) {
    if (fastq) {
      for (int f = 0; f < 2; ++f)
        launch_ing_parse_fastq(raw[f].data(), ls[f].data(), n_lines[f], hdr[f].data(), n_take[f], f, phase, replace_n, recs.data(), &scal[0], st);
    } else {
      launch_ing_parse_sam(raw[0].data(), ls[0].data(), n_lines[0], recs.data(), &scal[0], st);
    }
    launch_ing_scan_recs(recs.data(), m, blk4.data(), pre.data(), st);
    IngPublish pb{};
    pb.pre = pre.data(); pb.m = m; pb.final = final; pb.fastq = fastq ? 1 : 0; pb.phase = phase;
    for (int f = 0; f < 2; ++f) { pb.ls[f] = ls[f].data(); pb.hdr[f] = hdr[f].data(); pb.n_rec[f] = n_rec[f]; pb.n_bytes[f] = nb[f]; }
    pb.err = &scal[0]; pb.host = host;
    launch_ing_publish(pb, st);
    if (host[7] != ~0ull) { *err_index = host[7] >> 8; return (int)(host[7] & 0xff); }
    consumed[0] = host[5]; consumed[1] = host[6];
  } else {
    consumed[0] = final ? full[0] : 0; consumed[1] = (final && fastq) ? full[1] : 0;
  }
  if (final) { consumed[0] = full[0]; consumed[1] = fastq ? full[1] : 0; }
  const uint64_t n_reads = host[0], opt_bytes = host[3], m_used = host[4];
  out_info[0] = n_reads; out_info[1] = host[1]; out_info[2] = host[2]; out_info[3] = opt_bytes; out_info[4] = consumed[0]; out_info[5] = consumed[1]; out_info[6] = host[12];
  if (n_reads) {
    IngCopy cp{};
    cp.text[0] = raw[0].data(); cp.text[1] = raw[1].data(); cp.recs = recs.data(); cp.pre = pre.data(); cp.m = m_used;
    cp.names = names; cp.name_off = name_off; cp.seq = seq; cp.qual = qual; cp.seq_off = seq_off;
    cp.opt = opt_bytes ? opt : nullptr; cp.opt_off = opt_bytes ? opt_off : nullptr; cp.read_flag = read_flag;
    launch_ing_copy(cp, st);
    if (!opt_bytes) for (uint64_t i = 0; i <= n_reads; ++i) opt_off[i] = 0;     // smash_fetch_batch's convention
  } else {
    name_off[0] = seq_off[0] = opt_off[0] = 0;
  }
  return 0;
}
