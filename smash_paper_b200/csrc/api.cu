// api.cu -- host side of the C ABI (include/smash_b200.h): index files -> HBM, batch slots,
// stream plumbing.  No compute happens here and there is no CPU fallback: without a CUDA device
// every entry point that would compute returns SMASH_ERR_CUDA.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sched.h>

#include "../../include/smash_b200.h"
#include "kernels.cuh"
#include "tail.cuh"
#include "sabuild.cuh"
#include "ingest_launch.cuh"
#include "expand.h"
#include "ctx_internal.h"

using namespace smash;

static thread_local char g_err[1024] = "";
#include <chrono>
static const bool g_dbg = getenv("SMASH_DEBUG_TIMING") != nullptr;
static cudaEvent_t g_tl_base = nullptr;
static const bool g_force_split = getenv("SMASH_FORCE_SPLIT_SEARCH") != nullptr;
static const bool g_no_split = getenv("SMASH_NO_SPLIT_SEARCH") != nullptr;   // A/B switch: verification inside k_mam_search
static const bool g_no_tail_overlap = getenv("SMASH_NO_TAIL_OVERLAP") != nullptr;   // A/B switch: tail append after the emit kernels, same stream
static const bool g_no_chunks = getenv("SMASH_NO_CHUNKS") != nullptr;   // A/B switch for the chunked submit pipeline
static const bool g_full_sam = getenv("SMASH_FULL_SAM_D2H") != nullptr; // A/B switch: whole SAM text over PCIe instead of the compact transport
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define DBG_T(label, t0) do { if (g_dbg) fprintf(stderr, "[smash-dbg] %-28s %8.3f ms\n", label, now_ms() - (t0)); } while (0)
static int fail(int code, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return code;
}
#define CU(call)                                                                        \
  do { cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) return fail(SMASH_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

extern "C" const char *smash_last_error(void) { return g_err; }

extern "C" int smash_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major >= 10) ++ok;
  }
  return ok;
}

// ------------------------------------------------------------------ index (host view)

struct Mapping { void *p = nullptr; size_t n = 0; };
static int map_file(const std::string &path, size_t expect, Mapping *m) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return fail(SMASH_ERR_IO, "could not open input %s for reading", path.c_str());
  struct stat st; fstat(fd, &st);
  if (expect != (size_t)-1 && (size_t)st.st_size != expect) {
    close(fd);
    return fail(SMASH_ERR_IO, "%s has %zu bytes, expected %zu", path.c_str(), (size_t)st.st_size, expect);
  }
  m->n = (size_t)st.st_size;
  if (m->n) {
    m->p = mmap(nullptr, m->n, PROT_READ, MAP_SHARED, fd, 0);
    if (m->p == MAP_FAILED) { close(fd); m->p = nullptr; return fail(SMASH_ERR_IO, "Memory mapping error for %s", path.c_str()); }
  }
  close(fd);
  return 0;
}

struct smash_index {
  uint64_t N = 0; int w = 4; int rcref = 1; uint64_t fasta_size = 0;
  const uint8_t *text = nullptr; const void *sa = nullptr; const void *isa = nullptr;
  const uint8_t *lcp = nullptr; const uint8_t *lcp_m_raw = nullptr; uint64_t n_m = 0;
  std::vector<uint64_t> startpos, sizes; std::vector<std::string> descr;
  std::vector<Mapping> maps;
  std::string fasta;
};

static bool read_u64(FILE *f, uint64_t *v) { return fread(v, 8, 1, f) == 1; }

extern "C" int smash_index_open(const char *ref_fasta, int rcref, smash_index **out) {
  if (!ref_fasta || !out) return fail(SMASH_ERR_ARG, "null argument");
  smash_index *ix = new smash_index();
  ix->rcref = rcref ? 1 : 0; ix->fasta = ref_fasta;
  struct stat st;
  if (stat(ref_fasta, &st) != 0) { delete ix; return fail(SMASH_ERR_IO, "unable to open %s", ref_fasta); }
  ix->fasta_size = (uint64_t)st.st_size;
  const std::string base = std::string(ref_fasta) + ".bin/rc" + (rcref ? "1" : "0");
  FILE *f = fopen((base + ".ref.bin").c_str(), "rb");
  if (!f) { delete ix; return fail(SMASH_ERR_IO, "could not open reference bin file %s.ref.bin for reading (run the index build first)", base.c_str()); }
  uint64_t saved = 0, nd = 0;
  bool ok = read_u64(f, &saved) && read_u64(f, &ix->N) && read_u64(f, &nd);
  if (ok && saved != ix->fasta_size) {
    fclose(f); delete ix;
    return fail(SMASH_ERR_IO, "reference fasta size has changed\nmaybe the reference has changed?\nIf so, you will need to delete the current reference to proceed");
  }
  for (uint64_t i = 0; ok && i < nd; ++i) {
    uint64_t sp, sz, sl;
    ok = read_u64(f, &sp) && read_u64(f, &sz) && read_u64(f, &sl);
    if (!ok || sl > 1u << 20) { ok = false; break; }
    std::string d(sl, '\0');
    if (sl && fread(&d[0], 1, sl, f) != sl) { ok = false; break; }
    ix->startpos.push_back(sp); ix->sizes.push_back(sz); ix->descr.push_back(d);
  }
  fclose(f);
  if (!ok) { delete ix; return fail(SMASH_ERR_IO, "problem reading %s.ref.bin", base.c_str()); }
  // int width: the reference picks the binary (mummer / mummer-long) from the FASTA size
  // (mummer.cpp:156-183); we take whichever index the build left behind.
  std::string ib;
  if (access((base + ".i4.index.bin").c_str(), R_OK) == 0) { ix->w = 4; ib = base + ".i4.index"; }
  else if (access((base + ".i8.index.bin").c_str(), R_OK) == 0) { ix->w = 8; ib = base + ".i8.index"; }
  else { delete ix; return fail(SMASH_ERR_IO, "could not open index %s.i{4,8}.index.bin for reading", base.c_str()); }
  uint64_t hdr[6];
  f = fopen((ib + ".bin").c_str(), "rb");
  if (!f || fread(hdr, 8, 6, f) != 6) { if (f) fclose(f); delete ix; return fail(SMASH_ERR_IO, "problem reading 48 elements at index header"); }
  fclose(f);
  if (hdr[0] != ix->fasta_size) {
    delete ix;
    return fail(SMASH_ERR_IO, "saved fasta size used for index does notmatch current fasta size\nmaybe the reference has changed?\nyou may need to delete the current index to proceed");
  }
  if (hdr[3] != ix->N || hdr[4] != ix->N) { delete ix; return fail(SMASH_ERR_IO, "index size %llu does not match reference length %llu", (unsigned long long)hdr[3], (unsigned long long)ix->N); }
  ix->n_m = hdr[5];
  Mapping m;
  int rc;
  if ((rc = map_file(base + ".ref.seq.bin", ix->N, &m))) { delete ix; return rc; }
  ix->text = (const uint8_t *)m.p; ix->maps.push_back(m);
  if ((rc = map_file(ib + ".sa.bin", ix->N * ix->w, &m))) { smash_index_close(ix); return rc; }
  ix->sa = m.p; ix->maps.push_back(m);
  if (access((ib + ".isa.bin").c_str(), R_OK) == 0) {
    if ((rc = map_file(ib + ".isa.bin", ix->N * ix->w, &m))) { smash_index_close(ix); return rc; }
    ix->isa = m.p; ix->maps.push_back(m);
  }
  if ((rc = map_file(ib + ".lcp.vec.bin", ix->N, &m))) { smash_index_close(ix); return rc; }
  ix->lcp = (const uint8_t *)m.p; ix->maps.push_back(m);
  if ((rc = map_file(ib + ".lcp.m.bin", ix->n_m * 16, &m))) { smash_index_close(ix); return rc; }
  ix->lcp_m_raw = (const uint8_t *)m.p; ix->maps.push_back(m);
  *out = ix;
  return 0;
}

extern "C" int smash_index_from_arrays(const uint8_t *text, uint64_t N, const void *sa, const void *isa, int w,
                                       const uint8_t *lcp_vec, const void *lcp_m, uint64_t n_m,
                                       uint64_t n_descr, const uint64_t *startpos, const uint64_t *sizes,
                                       const char *const *descr, int rcref, smash_index **out) {
  if (!text || !sa || !lcp_vec || !out || (w != 4 && w != 8) || !n_descr) return fail(SMASH_ERR_ARG, "bad index arrays");
  smash_index *ix = new smash_index();
  ix->N = N; ix->w = w; ix->rcref = rcref ? 1 : 0; ix->text = text; ix->sa = sa; ix->isa = isa;
  ix->lcp = lcp_vec; ix->lcp_m_raw = (const uint8_t *)lcp_m; ix->n_m = n_m;
  for (uint64_t i = 0; i < n_descr; ++i) { ix->startpos.push_back(startpos[i]); ix->sizes.push_back(sizes[i]); ix->descr.push_back(descr[i]); }
  *out = ix;
  return 0;
}

extern "C" void smash_index_close(smash_index *ix) {
  if (!ix) return;
  for (auto &m : ix->maps) if (m.p && m.n) munmap(m.p, m.n);
  delete ix;
}
extern "C" uint64_t smash_index_text_len(const smash_index *ix) { return ix ? ix->N : 0; }
extern "C" int smash_index_int_width(const smash_index *ix) { return ix ? ix->w : 0; }

extern "C" size_t smash_index_sam_header(const smash_index *ix, char *buf, size_t cap) {
  std::string s = "@HD\tVN:1.0\tSO:unsorted\n";                 // fasta.cpp:243-252
  for (size_t c = 0; c < ix->sizes.size(); c += ix->rcref ? 2 : 1)
    s += "@SQ\tSN:" + ix->descr[c] + "\tLN:" + std::to_string(ix->sizes[c]) + "\n";
  s += "@PG\tID:longMEM\tPN:longMEM\tVN:0.5\n";
  if (buf && cap) { size_t n = s.size() < cap ? s.size() : cap; memcpy(buf, s.data(), n); }
  return s.size();
}

// ------------------------------------------------------------------ context

template <class T> struct DBuf {        // growable device buffer
  T *p = nullptr; size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = n + n / 4 + 64;
    cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
    if (e != cudaSuccess) { cap = 0; return fail(SMASH_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e)); }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T> struct HBuf {        // growable pinned host buffer
  T *p = nullptr; size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    size_t want = n + n / 4 + 64;
    cudaError_t e = cudaHostAlloc((void **)&p, want * sizeof(T), cudaHostAllocMapped);
    if (e != cudaSuccess) { cap = 0; return fail(SMASH_ERR_NOMEM, "cudaHostAlloc(%zu bytes): %s", want * sizeof(T), cudaGetErrorString(e)); }
    cap = want;
    return 0;
  }
  int grow_keep(size_t n, size_t keep) {        // like ensure(), but the first `keep` elements survive
    if (n <= cap) return 0;
    T *old = p; p = nullptr; cap = 0;
    int rc = ensure(n);
    if (!rc && old && keep) memcpy(p, old, keep * sizeof(T));
    if (old) cudaFreeHost(old);
    return rc;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// A submitted batch is cut into up to MAX_CHUNKS read ranges that flow through three streams of the
// slot (upload -> kernels -> download), so that the SAM text of the first range is already crossing
// PCIe while the later ranges are still being searched.
constexpr int MAX_CHUNKS = 8;
constexpr uint64_t CHUNK_MIN_READS = 65536;
constexpr int N_EVS = 10 * MAX_CHUNKS;

struct Slot {
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool owns_out = false;
  cudaStream_t st_in = nullptr, st_out = nullptr;                  // chunked submit: upload / download streams
  cudaStream_t st_tail = nullptr; cudaEvent_t ev_tail = nullptr;   // tail append of a range next to its emit kernels
  cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_emit[MAX_CHUNKS] = {}, ev_out = nullptr;
  int n_chunks = 1; uint64_t chunk_r[MAX_CHUNKS + 1] = {}; uint64_t sam_base = 0;
  uint64_t chunk_name_bytes[MAX_CHUNKS] = {};                      // name bytes of each read range (upper bound)
  cudaEvent_t evs[N_EVS] = {};   // stage boundaries
  int n_evs = 0; int ev_stage[N_EVS];
  cudaEvent_t tl[3] = {nullptr, nullptr, nullptr};   // SMASH_DEBUG_TIMING timeline: submit, D2H begin, D2H end
  // batch on device
  DBuf<uint8_t> names, seq, qual, opt;
  DBuf<int64_t> name_off, seq_off, opt_off;
  DBuf<uint16_t> read_flag;
  BatchDev bd{};
  // work
  int cap = 0;
  DBuf<Match> match_slots; DBuf<Match> mem_stage; DBuf<uint32_t> match_cnt; DBuf<Item> item_slots; DBuf<Rec> rec_slots;
  DBuf<ReadSum> sums; DBuf<uint32_t> nrec; DBuf<uint64_t> rec_base; DBuf<uint32_t> rec_read; DBuf<uint32_t> rec_bytes; DBuf<uint64_t> rec_off; DBuf<uint64_t> sam_total; DBuf<uint64_t> blk_sums2; DBuf<uint64_t> blk_sums;
  DBuf<char> sam; DBuf<uint32_t> flags;
  DBuf<int64_t> csr_off; DBuf<uint64_t> csr_triples;
  DBuf<uint64_t> slot_off; DBuf<Aln> aln_scr; DBuf<uint16_t> ord_scr; DBuf<uint32_t> tmp32;   // MEM mode (CSR slots)
  uint64_t slots_total = 0; bool csr = false;
  DBuf<uint8_t> long_scratch; int long_q = 0;
  bool split = false, seed_ok = false;
  DBuf<uint64_t> surv; DBuf<uint8_t> surv_cnt; DBuf<uint8_t> lc; DBuf<uint8_t> slow;      // split search (k_mam_seed [-> k_mam_search] -> k_mam_verify)
  DBuf<uint64_t> sort_abs, sort_off; DBuf<uint8_t> sort_flag, sort_tmp; DBuf<uint32_t> sort_perm, sort_bytes;   // K5 record_sort
  bool sorted = false;
  HBuf<uint16_t> h_flag;       // pinned staging for a pageable read_flag array
  // device-side input stage (smash_submit_text): raw text, line tables, per-record parse results, scan scratch
  DBuf<uint8_t> ing_raw[2]; DBuf<uint64_t> ing_ls[2]; DBuf<uint64_t> ing_hdr[2]; DBuf<uint8_t> ing_hdr_flag;
  DBuf<uint64_t> ing_blk64; DBuf<uint32_t> ing_blk32; DBuf<Ing4> ing_blk4; DBuf<Ing4> ing_pre; DBuf<LineRec> ing_recs;
  DBuf<unsigned long long> ing_scal;      // [0] first error, [1..2] FASTQ records per text
  HBuf<uint64_t> h_ing;
  uint64_t name_bytes = 0, seq_bytes = 0, opt_bytes = 0;   // sizes of the batch blobs on the device
  cudaEvent_t ev_ing0 = nullptr, ev_ing1 = nullptr; bool ing_timed = false;   // device time of the input stage (after the H2D copy)
  // results on host
  HBuf<char> h_sam; HBuf<int64_t> h_csr_off; HBuf<smash_match> h_matches; HBuf<uint64_t> h_small;
  // in flight
  bool busy = false; int want = 0; uint64_t n_reads = 0; uint64_t first_pair = 0;
  uint64_t sam_bytes = 0, n_matches = 0, n_records = 0;
  uint64_t launches = 0, io_h2d = 0, io_d2h = 0;                  // merged into the ctx by slot_finish (caller's thread)
  // compact transport (compact.h): per read range, head|tags|lr text + CmpMeta cross PCIe; host threads rebuild the lines
  bool compact = false;
  DBuf<uint32_t> cmp_bytes; DBuf<uint64_t> cmp_off;
  DBuf<char> cmp[MAX_CHUNKS]; DBuf<CmpMeta> cmeta[MAX_CHUNKS];
  HBuf<char> h_cmp[MAX_CHUNKS]; HBuf<CmpMeta> h_cmeta[MAX_CHUNKS];
  cudaEvent_t ev_d2h[MAX_CHUNKS] = {};
  uint64_t cmp_recs[MAX_CHUNKS] = {}; int n_ranges = 0, n_dispatched = 0;
  smash_ctx *ctx = nullptr;
  std::atomic<int> pending{0};                                     // expansion tasks not finished yet
  std::atomic<uint64_t> expand_ns{0};                              // SMASH_DEBUG_TIMING: summed task time of the batch
  // the slot's host worker: H2D enqueue, the per-range launches (with their size read-backs) and the expansion run
  // here, so smash_submit returns at once and two slots never serialise on each other's host synchronisations
  std::thread worker; std::mutex mu; std::condition_variable cv;
  int job_state = 0;                                               // 0 idle, 1 queued, 2 running, 3 done
  bool quit = false;
  smash_batch hb{}; bool job_prepare = false; int job_want = 0, job_chunks = 1, job_rc = 0; char job_err[1024] = "";
  uint64_t job_seq = 0; bool turn_held = false, turn_done = true;     // order of the tail appends = order of submission
};

// host threads that expand compact records into SAM lines (one pool per context)
struct ExpandTask { Slot *s; int ch; uint64_t f0, f1; };
struct HostPool {
  std::vector<std::thread> th; std::mutex mu; std::condition_variable cv; std::deque<ExpandTask> q; bool quit = false;
};

struct smash_ctx {
  const smash_index *hix = nullptr;
  smash_params prm{};
  int device = 0;
  DevIndex dix{};
  SearchParams sp{};
  // index storage
  uint8_t *text_alloc = nullptr; void *sa = nullptr; void *isa = nullptr; uint8_t *lcp = nullptr;
  LcpItem *lcp_m = nullptr; uint8_t *uniq = nullptr; void *seed = nullptr; uint64_t *seed_irr = nullptr; uint32_t *ext = nullptr; uint64_t *startpos = nullptr;
  uint64_t *sizes = nullptr; char *descr = nullptr; int *descr_off = nullptr; uint64_t *descr8 = nullptr; uint32_t *alpha = nullptr;
  uint8_t *mapbody = nullptr; uint32_t *chrom_off32 = nullptr; uint64_t *chrom_abs64 = nullptr;
  uint64_t n_m = 0;
  int max_chunks = MAX_CHUNKS; uint64_t chunk_min_reads = CHUNK_MIN_READS;
  smash_index *own_index = nullptr;
  uint64_t index_bytes = 0;
  uint64_t launches = 0, io_h2d = 0, io_d2h = 0;
  int transport = 0;                                // 0: per read range, whichever is faster; 1: full SAM text over PCIe; 2: compact only
  int host_threads = 0;                             // expansion threads (0 = auto)
  void *comm = nullptr;                              // multi-GPU state (comm.cu)
  HostPool pool;
  // transport scheduler: when the download stream and the line-building threads are expected to be free [host clock, ms],
  // the measured rate of each; a read range takes whichever way gets its lines into host memory first
  std::mutex sched_mu; double dma_free_at = 0, cpu_free_at = 0, dma_bytes_per_ms = 45e6, cpu_ns_per_rec = 100.0; bool dma_calibrated = false;
  std::mutex turn_mu; std::condition_variable turn_cv; uint64_t next_seq = 0, tail_turn = 0;
  double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // search, records, sizes+scan, emit_text, csr, tail, emit_copy, verify
  double ingest_ms = 0;                            // device-side input stage (smash_submit_text / smash_text_upload)
  Slot slot[SMASH_N_SLOTS];
  TailState tail;
};

static int job_start(smash_ctx *c, Slot &s);

static int dmalloc(void **p, size_t bytes, uint64_t *acct) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) return fail(SMASH_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
  if (acct) *acct += bytes;
  return 0;
}

extern "C" void smash_params_default(smash_params *p) {
  memset(p, 0, sizeof *p);
  p->device = 0; p->mode = SMASH_MODE_MAM; p->min_len = 20; p->nomap = 0; p->nucleotides_only = 0;
  p->tag_mappability = 0; p->max_batch_reads = 0; p->seed_k = 0;
}

extern "C" void *smash_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void smash_host_free(void *p) { if (p) cudaFreeHost(p); }

static void set_search_params(smash_ctx *c) {
  SearchParams &sp = c->sp;
  sp.L = c->prm.min_len < 2 ? 2u : c->prm.min_len;
  sp.k = c->dix.seed_k < (int)sp.L ? c->dix.seed_k : (int)sp.L;
  sp.s = (int)sp.L - sp.k + 1;
  sp.nucleotides_only = c->prm.nucleotides_only;
  sp.nomap = c->prm.nomap;
  sp.tag_mappability = c->prm.tag_mappability;
  sp.mum = c->prm.mode == SMASH_MODE_MUM;
  // the anchor path enumerates seed buckets: only sensible when a k-mer of that length is rare
  const double expect = (double)c->dix.N / pow(4.0, (double)sp.k);
  sp.fast_ok = expect <= 16.0;
}

#define CK(x) do { if ((rc = (x))) { smash_ctx_destroy(c); return rc; } } while (0)
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { smash_ctx_destroy(c); return fail(SMASH_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)

static int ctx_begin(const smash_params *p, smash_ctx **out, int rcref) {
  if (p->tag_mappability && !rcref) return fail(SMASH_ERR_ARG, "tag_mappability requires an -rcref index (mummer.cpp:145)");
  if (smash_device_count() <= 0) return fail(SMASH_ERR_CUDA, "no sm_100 CUDA device available (this library has no CPU fallback)");
  if (p->mode != SMASH_MODE_MAM && p->mode != SMASH_MODE_MEM && p->mode != SMASH_MODE_MUM) return fail(SMASH_ERR_ARG, "mode %d not supported", p->mode);
  CU(cudaSetDevice(p->device));
  {
    // The index is probed with isolated 2..16-byte reads scattered over > 100 GB: ask the L2 to fetch single 32-byte
    // sectors on a miss instead of its default multi-sector granule (a hint; SMASH_L2_FETCH=64|128 for A/B runs).
    size_t gran = 32;
    if (const char *e = getenv("SMASH_L2_FETCH")) gran = (size_t)atoi(e);
    if (gran == 32 || gran == 64 || gran == 128) { if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) != cudaSuccess) cudaGetLastError(); }
  }
  smash_ctx *c = new smash_ctx();
  c->prm = *p; c->device = p->device;
  for (int s = 0; s < SMASH_N_SLOTS; ++s) {
    CUC(cudaStreamCreateWithFlags(&c->slot[s].st, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&c->slot[s].ev0)); CUC(cudaEventCreate(&c->slot[s].ev1));
    for (int e = 0; e < N_EVS; ++e) CUC(cudaEventCreate(&c->slot[s].evs[e]));
    CUC(cudaStreamCreateWithFlags(&c->slot[s].st_in, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->slot[s].st_tail, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&c->slot[s].ev_tail, cudaEventDisableTiming));
    if (s == 0) { CUC(cudaStreamCreateWithFlags(&c->slot[s].st_out, cudaStreamNonBlocking)); c->slot[s].owns_out = true; }
    else c->slot[s].st_out = c->slot[0].st_out;     // ONE download stream: SAM ranges cross PCIe one after another, in submit order
    for (int e = 0; e < MAX_CHUNKS; ++e) {
      CUC(cudaEventCreateWithFlags(&c->slot[s].ev_in[e], cudaEventDisableTiming));
      CUC(cudaEventCreateWithFlags(&c->slot[s].ev_emit[e], cudaEventDisableTiming));
      CUC(cudaEventCreateWithFlags(&c->slot[s].ev_d2h[e], cudaEventDisableTiming));
    }
    CUC(cudaEventCreateWithFlags(&c->slot[s].ev_out, cudaEventDisableTiming));
  }
  *out = c;
  return 0;
}

// metadata + derived structures once text/sa/(isa)/lcp/lcp_m are in HBM
static int ctx_finish(smash_ctx *c, const smash_index *ix) {
  int rc = 0;
  cudaStream_t st = c->slot[0].st;
  const uint64_t N = ix->N;
  const int nd = (int)ix->descr.size();
  CK(dmalloc((void **)&c->startpos, 8 * nd, &c->index_bytes));
  CK(dmalloc((void **)&c->sizes, 8 * nd, &c->index_bytes));
  CUC(cudaMemcpy(c->startpos, ix->startpos.data(), 8 * nd, cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(c->sizes, ix->sizes.data(), 8 * nd, cudaMemcpyHostToDevice));
  {
    std::string blob; std::vector<int> off(nd + 1, 0);
    for (int i = 0; i < nd; ++i) { off[i] = (int)blob.size(); blob += ix->descr[i]; }
    off[nd] = (int)blob.size();
    CK(dmalloc((void **)&c->descr, blob.size() + 1, &c->index_bytes));
    CK(dmalloc((void **)&c->descr_off, 4 * (nd + 1), &c->index_bytes));
    CUC(cudaMemcpy(c->descr, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    CUC(cudaMemcpy(c->descr_off, off.data(), 4 * (nd + 1), cudaMemcpyHostToDevice));
    std::vector<uint64_t> d8(nd, 0);
    for (int i = 0; i < nd; ++i) for (size_t j = 0; j < ix->descr[i].size() && j < 8; ++j) d8[i] |= (uint64_t)(uint8_t)ix->descr[i][j] << (8 * j);
    CK(dmalloc((void **)&c->descr8, 8 * nd, &c->index_bytes));
    CUC(cudaMemcpy(c->descr8, d8.data(), 8 * nd, cudaMemcpyHostToDevice));
    // mappability_tag's 32-bit offsets over all @SQ (chromosomes.h:29-66)
    std::vector<uint32_t> o32((nd + 1) / (ix->rcref ? 2 : 1) + 1, 0);
    uint32_t acc = 0; int k = 0;
    for (int i = 0; i < nd; i += ix->rcref ? 2 : 1) { o32[k++] = acc; acc += (uint32_t)ix->sizes[i]; }
    CK(dmalloc((void **)&c->chrom_off32, 4 * o32.size(), &c->index_bytes));
    CUC(cudaMemcpy(c->chrom_off32, o32.data(), 4 * o32.size(), cudaMemcpyHostToDevice));
    // MemSam::chromosomes (query.cpp:546-552): 64-bit offsets of the forward sequences, "*" = total
    std::vector<uint64_t> o64;
    uint64_t acc64 = 0;
    for (int i = 0; i < nd; i += ix->rcref ? 2 : 1) { o64.push_back(acc64); acc64 += ix->sizes[i]; }
    o64.push_back(acc64);
    CK(dmalloc((void **)&c->chrom_abs64, 8 * o64.size(), &c->index_bytes));
    CUC(cudaMemcpy(c->chrom_abs64, o64.data(), 8 * o64.size(), cudaMemcpyHostToDevice));
  }
  CK(dmalloc((void **)&c->alpha, 32, nullptr));
  DevIndex &d = c->dix;
  d.text = c->text_alloc + TEXT_PAD; d.N = N; d.sa = c->sa; d.isa = c->isa; d.w = ix->w; d.lcp = c->lcp;
  d.lcp_m = c->lcp_m; d.n_m = c->n_m; d.startpos = c->startpos; d.sizes = c->sizes; d.n_descr = nd;
  d.rcref = ix->rcref; d.descr = c->descr; d.descr_off = c->descr_off; d.descr8 = c->descr8;
  d.logN = (uint64_t)ceil(log((double)N) / log(2.0));          // longSA.cpp:97
  d.mapbody = nullptr; d.map_bytes = 0; d.chrom_off32 = c->chrom_off32; d.chrom_abs64 = c->chrom_abs64;
  // derived: alphabet bitmap, shortest-unique-length bytes, k-mer seed table
  c->launches += launch_alpha(d.text, N, c->alpha, st);
  CUC(cudaMemcpyAsync(d.alpha, c->alpha, 32, cudaMemcpyDeviceToHost, st));
  CK(dmalloc((void **)&c->uniq, N, &c->index_bytes));
  c->launches += launch_uniq_build(d, c->uniq, st);
  d.uniq = c->uniq;
  int k = c->prm.seed_k;
  if (k <= 0) { k = (int)ceil(log((double)N) / log(4.0)) + 1; }
  if (k > 16) k = 16;
  if (k < 4) k = 4;
  // seed entries follow the index's integer width (size.h:9-22): an i8 index of a small text runs the very
  // code path of the hg19-scale one (8-byte SA entries AND 8-byte seed entries)
  d.seed_k = k; d.seed_w = (ix->w == 8 || N >= 0xffffffffull) ? 8 : 4;
  CK(dmalloc(&c->seed, ((1ull << (2 * k)) + 1) * d.seed_w, &c->index_bytes));
  c->launches += launch_seed_build(d, c->seed, k, d.seed_w, st);
  d.seed = c->seed;
  set_search_params(c);
  d.ext = nullptr;
  if (c->prm.mode != SMASH_MODE_MEM && c->sp.fast_ok && (double)N / pow(4.0, (double)c->sp.k) >= 0.25) {
    // the 8+6 character pre-filter pays off when chance hits of the seed are common (large references)
    CK(dmalloc((void **)&c->ext, 4 * N, &c->index_bytes));
    c->launches += launch_ext_build(d, c->sp.k, c->ext, st);
    d.ext = c->ext;
    // blocked seed table (core.cuh): the 8-byte table of an index whose anchors use the full seed length is rewritten in
    // place, so that a bucket's bounds and its candidates' ext codes share one 128-byte line.  SMASH_FLAT_SEED: A/B switch
    if (d.seed_w == 8 && c->sp.k == d.seed_k && k >= 2 && !getenv("SMASH_FLAT_SEED")) {
      const uint64_t n_buckets = 1ull << (2 * k), n_blocks = n_buckets >> 4;
      unsigned long long *d_cnt = nullptr, h_cnt = 0;
      uint64_t h_end = 0;
      CK(dmalloc((void **)&d_cnt, 8, nullptr));
      CUC(cudaMemsetAsync(d_cnt, 0, 8, st));
      c->launches += launch_seed_irr_count(c->seed, n_blocks, d_cnt, st);
      CUC(cudaMemcpyAsync(&h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st));
      CUC(cudaMemcpyAsync(&h_end, (const uint64_t *)c->seed + n_buckets, 8, cudaMemcpyDeviceToHost, st));
      CUC(cudaStreamSynchronize(st));
      CK(dmalloc((void **)&c->seed_irr, 8 * 17 * (size_t)(h_cnt + 1), &c->index_bytes));
      CUC(cudaMemsetAsync(d_cnt, 0, 8, st));
      c->launches += launch_seed_block(c->seed, n_blocks, c->ext, N, c->seed_irr, d_cnt, st);
      CUC(cudaStreamSynchronize(st));
      cudaFree(d_cnt);
      d.seed_blocked = 1; d.seed_n = n_buckets; d.seed_end = h_end; d.seed_irr = c->seed_irr;
      if (g_dbg) fprintf(stderr, "[smash-dbg] blocked seed table: %llu blocks of 16 buckets, %llu with a row of exact values\n", (unsigned long long)n_blocks, h_cnt);
    }
  }
  CUC(cudaStreamSynchronize(st));
  CUC(cudaGetLastError());
  tail_init(&c->tail);
  return 0;
}

extern "C" int smash_ctx_create(const smash_index *ix, const smash_params *p, smash_ctx **out) {
  if (!ix || !p || !out) return fail(SMASH_ERR_ARG, "null argument");
  if (!ix->text || !ix->sa) return fail(SMASH_ERR_ARG, "index has no host arrays");
  if (p->mode == SMASH_MODE_MEM && !ix->isa) return fail(SMASH_ERR_ARG, "MEM mode needs the .isa.bin array");
  smash_ctx *c = nullptr;
  int rc = ctx_begin(p, &c, ix->rcref);
  if (rc) return rc;
  c->hix = ix;
  const uint64_t N = ix->N; const int w = ix->w;
  cudaStream_t st = c->slot[0].st;
  CK(dmalloc((void **)&c->text_alloc, N + 2 * TEXT_PAD, &c->index_bytes));
  CUC(cudaMemsetAsync(c->text_alloc, 0, N + 2 * TEXT_PAD, st));
  CUC(cudaMemcpyAsync(c->text_alloc + TEXT_PAD, ix->text, N, cudaMemcpyHostToDevice, st));
  CK(dmalloc(&c->sa, N * w, &c->index_bytes));
  CUC(cudaMemcpyAsync(c->sa, ix->sa, N * w, cudaMemcpyHostToDevice, st));
  if (ix->isa && (p->mode == SMASH_MODE_MEM)) {
    CK(dmalloc(&c->isa, N * w, &c->index_bytes));
    CUC(cudaMemcpyAsync(c->isa, ix->isa, N * w, cudaMemcpyHostToDevice, st));
  }
  CK(dmalloc((void **)&c->lcp, N, &c->index_bytes));
  CUC(cudaMemcpyAsync(c->lcp, ix->lcp, N, cudaMemcpyHostToDevice, st));
  {
    // .lcp.m.bin items are {u64 idx; ANINT val; pad}: widen val so the device sees one layout
    std::vector<LcpItem> m(ix->n_m);
    for (uint64_t i = 0; i < ix->n_m; ++i) {
      const uint8_t *r = ix->lcp_m_raw + 16 * i;
      memcpy(&m[i].idx, r, 8);
      if (w == 4) { uint32_t v; memcpy(&v, r + 8, 4); m[i].val = v; } else memcpy(&m[i].val, r + 8, 8);
    }
    CK(dmalloc((void **)&c->lcp_m, sizeof(LcpItem) * (ix->n_m + 1), &c->index_bytes));
    if (ix->n_m) CUC(cudaMemcpy(c->lcp_m, m.data(), sizeof(LcpItem) * ix->n_m, cudaMemcpyHostToDevice));
    c->n_m = ix->n_m;
  }
  if ((rc = ctx_finish(c, ix))) return rc;
  *out = c;
  return 0;
}

// Index construction on the GPU (replaces the longSA build branch, longSA.cpp:137-176): the text
// is uploaded, SA/ISA/LCP are built in HBM and stay there; nothing is copied back unless
// smash_ctx_save_index / smash_ctx_copy_index is called.
extern "C" int smash_ctx_create_from_text(const uint8_t *text, uint64_t N, uint64_t n_descr, const uint64_t *startpos,
                                          const uint64_t *sizes, const char *const *descr, int rcref, int w,
                                          int keep_isa, uint64_t chunk_cap, const smash_params *p, smash_ctx **out) {
  if (!text || !N || !n_descr || !startpos || !sizes || !descr || !p || !out || (w != 4 && w != 8))
    return fail(SMASH_ERR_ARG, "bad argument");
  if (w == 4 && N >= 0xffffffffull) return fail(SMASH_ERR_ARG, "text too long for 4-byte index integers");
  smash_ctx *c = nullptr;
  int rc = ctx_begin(p, &c, rcref ? 1 : 0);
  if (rc) return rc;
  smash_index *ix = new smash_index();
  ix->N = N; ix->w = w; ix->rcref = rcref ? 1 : 0;
  for (uint64_t i = 0; i < n_descr; ++i) { ix->startpos.push_back(startpos[i]); ix->sizes.push_back(sizes[i]); ix->descr.push_back(descr[i]); }
  c->hix = ix; c->own_index = ix;
  cudaStream_t st = c->slot[0].st;
  CK(dmalloc((void **)&c->text_alloc, N + 2 * TEXT_PAD, &c->index_bytes));
  CUC(cudaMemsetAsync(c->text_alloc, 0, N + 2 * TEXT_PAD, st));
  CUC(cudaMemcpyAsync(c->text_alloc + TEXT_PAD, text, N, cudaMemcpyHostToDevice, st));
  CK(dmalloc(&c->sa, N * w, &c->index_bytes));
  if (keep_isa || p->mode == SMASH_MODE_MEM) CK(dmalloc(&c->isa, N * w, &c->index_bytes));
  CK(dmalloc((void **)&c->lcp, N, &c->index_bytes));
  char err[256] = "";
  rc = build_index_device(c->text_alloc + TEXT_PAD, N, w, c->sa, c->isa, c->lcp, &c->lcp_m, &c->n_m, chunk_cap, st, err, &c->launches);
  if (rc) { smash_ctx_destroy(c); return fail(rc == -3 ? SMASH_ERR_CUDA : SMASH_ERR_DATA, "index build: %s", err); }
  c->index_bytes += sizeof(LcpItem) * (c->n_m + 1);
  if ((rc = ctx_finish(c, ix))) return rc;
  *out = c;
  return 0;
}
#undef CK
#undef CUC

// Copy the index arrays from HBM to caller buffers (any pointer may be NULL).  lcp_m receives
// n_m 16-byte items in the .lcp.m.bin layout {u64 idx; ANINT val; zero pad}.
extern "C" int smash_ctx_copy_index(smash_ctx *c, void *sa, void *isa, uint8_t *lcp_vec, void *lcp_m, uint64_t *n_m) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  const uint64_t N = c->dix.N; const int w = c->dix.w;
  if (sa) CU(cudaMemcpy(sa, c->sa, N * w, cudaMemcpyDeviceToHost));
  if (isa) { if (!c->isa) return fail(SMASH_ERR_STATE, "ISA is not resident"); CU(cudaMemcpy(isa, c->isa, N * w, cudaMemcpyDeviceToHost)); }
  if (lcp_vec) CU(cudaMemcpy(lcp_vec, c->lcp, N, cudaMemcpyDeviceToHost));
  if (n_m) *n_m = c->n_m;
  if (lcp_m && c->n_m) {
    std::vector<LcpItem> m(c->n_m);
    CU(cudaMemcpy(m.data(), c->lcp_m, sizeof(LcpItem) * c->n_m, cudaMemcpyDeviceToHost));
    uint8_t *o = (uint8_t *)lcp_m;
    for (uint64_t i = 0; i < c->n_m; ++i) {
      memset(o + 16 * i, 0, 16); memcpy(o + 16 * i, &m[i].idx, 8);
      if (w == 4) { uint32_t v = (uint32_t)m[i].val; memcpy(o + 16 * i + 8, &v, 4); } else memcpy(o + 16 * i + 8, &m[i].val, 8);
    }
  }
  return 0;
}

static int write_file(const std::string &path, const void *p, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return fail(SMASH_ERR_IO, "could not open output %s for writing", path.c_str());
  const bool ok = n == 0 || fwrite(p, 1, n, f) == n;
  if (fclose(f) != 0 || !ok) return fail(SMASH_ERR_IO, "problem writing %zu bytes at %s", n, path.c_str());
  return 0;
}
static int write_device_file(const std::string &path, const void *dptr, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return fail(SMASH_ERR_IO, "could not open output %s for writing", path.c_str());
  const size_t CH = 256u << 20;
  void *h = nullptr;
  if (cudaHostAlloc(&h, CH, cudaHostAllocDefault) != cudaSuccess) { fclose(f); return fail(SMASH_ERR_NOMEM, "cudaHostAlloc"); }
  int rc = 0;
  for (size_t o = 0; o < n && !rc; o += CH) {
    const size_t m = n - o < CH ? n - o : CH;
    if (cudaMemcpy(h, (const uint8_t *)dptr + o, m, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(SMASH_ERR_CUDA, "D2H copy failed");
    else if (fwrite(h, 1, m, f) != m) rc = fail(SMASH_ERR_IO, "problem writing %s", path.c_str());
  }
  cudaFreeHost(h);
  if (fclose(f) != 0 && !rc) rc = fail(SMASH_ERR_IO, "problem closing %s", path.c_str());
  return rc;
}
// Write <fa>.bin/rc{r}.* exactly as the reference's build branch does (fasta.cpp:215-236,
// longSA.cpp:179-190) from the arrays in HBM.  Needs ISA resident.
extern "C" int smash_ctx_save_index(smash_ctx *c, const char *ref_fasta, int with_mappability) {
  if (!c || !ref_fasta) return fail(SMASH_ERR_ARG, "null argument");
  if (!c->isa) return fail(SMASH_ERR_STATE, "ISA is not resident (create the context with keep_isa)");
  CU(cudaSetDevice(c->device));
  struct stat st;
  if (stat(ref_fasta, &st) != 0) return fail(SMASH_ERR_IO, "unable to open %s", ref_fasta);
  const uint64_t fasta_size = (uint64_t)st.st_size;
  const smash_index *ix = c->hix;
  const std::string dir = std::string(ref_fasta) + ".bin";
  mkdir(dir.c_str(), 0755);
  const std::string base = dir + "/rc" + (ix->rcref ? "1" : "0");
  std::string ref;
  auto put64 = [&](uint64_t v) { ref.append((const char *)&v, 8); };
  put64(fasta_size); put64(ix->N); put64(ix->descr.size());
  uint64_t maxd = 0;
  for (size_t i = 0; i < ix->descr.size(); ++i) {
    put64(ix->startpos[i]); put64(ix->sizes[i]); put64(ix->descr[i].size()); ref += ix->descr[i];
    if (ix->descr[i].size() > maxd) maxd = ix->descr[i].size();
  }
  put64(maxd);
  int rc;
  if ((rc = write_file(base + ".ref.bin", ref.data(), ref.size()))) return rc;
  if ((rc = write_device_file(base + ".ref.seq.bin", c->text_alloc + TEXT_PAD, ix->N))) return rc;
  const std::string ib = base + ".i" + std::to_string(ix->w) + ".index";
  const uint64_t hdr[6] = {fasta_size, c->dix.logN, ix->N - 1, ix->N, ix->N, c->n_m};
  if ((rc = write_file(ib + ".bin", hdr, 48))) return rc;
  if ((rc = write_device_file(ib + ".sa.bin", c->sa, ix->N * ix->w))) return rc;
  if ((rc = write_device_file(ib + ".isa.bin", c->isa, ix->N * ix->w))) return rc;
  if ((rc = write_device_file(ib + ".lcp.vec.bin", c->lcp, ix->N))) return rc;
  std::vector<uint8_t> m(16 * c->n_m);
  if ((rc = smash_ctx_copy_index(c, nullptr, nullptr, nullptr, m.data(), nullptr))) return rc;
  if ((rc = write_file(ib + ".lcp.m.bin", m.data(), m.size()))) return rc;
  if (with_mappability) {
    if (!c->mapbody) return fail(SMASH_ERR_STATE, "mappability has not been built");
    FILE *f = fopen((dir + "/map.bin").c_str(), "wb");
    if (!f) return fail(SMASH_ERR_IO, "could not open outfile %s/map.bin for writing", dir.c_str());
    fputc(0, f); fputc(0, f); fclose(f);          // the reference writes two junk bytes first (longSA.cpp:606-617)
    std::vector<uint8_t> body(c->dix.map_bytes);
    CU(cudaMemcpy(body.data(), c->mapbody, body.size(), cudaMemcpyDeviceToHost));
    f = fopen((dir + "/map.bin").c_str(), "ab");
    if (!f || fwrite(body.data(), 1, body.size(), f) != body.size()) { if (f) fclose(f); return fail(SMASH_ERR_IO, "problem writing map.bin"); }
    fclose(f);
  }
  return 0;
}

// ------------------------------------------------------------------ host threads: expansion pool, tail turn

static void sched_note_task(smash_ctx *c, double ns_per_rec);
static void sched_calibrate_dma(smash_ctx *c, cudaStream_t st);
static void expand_task_run(const ExpandTask &t) {
  Slot &s = *t.s;
  ExpandArgs a{};
  a.names = s.hb.names; a.name_off = s.hb.name_off; a.seq = s.hb.seq; a.qual = s.hb.qual; a.seq_off = s.hb.seq_off;
  a.opt = s.opt_bytes ? s.hb.opt : nullptr; a.opt_off = s.opt_bytes ? s.hb.opt_off : nullptr;
  a.read_base = s.chunk_r[t.ch];
  a.meta = s.h_cmeta[t.ch].p; a.cmp = s.h_cmp[t.ch].p; a.sam = s.h_sam.p;
  const double te = now_ms();
  expand_records(a, t.f0, t.f1);
  const double dt_ns = (now_ms() - te) * 1e6;
  if (g_dbg) s.expand_ns.fetch_add((uint64_t)dt_ns);
  if (s.ctx && t.f1 > t.f0) sched_note_task(s.ctx, dt_ns / (double)(t.f1 - t.f0));
  if (s.pending.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(s.mu); s.cv.notify_all(); }
}
static bool pool_try_run_one(HostPool &p) {
  ExpandTask t;
  {
    std::lock_guard<std::mutex> lk(p.mu);
    if (p.q.empty()) return false;
    t = p.q.front(); p.q.pop_front();
  }
  expand_task_run(t);
  return true;
}
static void pool_main(HostPool *p) {
  for (;;) {
    ExpandTask t;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv.wait(lk, [&] { return p->quit || !p->q.empty(); });
      if (p->q.empty()) return;                              // quit
      t = p->q.front(); p->q.pop_front();
    }
    expand_task_run(t);
  }
}
static int host_cpu_count() {
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) { const int n = CPU_COUNT(&set); if (n > 0) return n; }
  const unsigned h = std::thread::hardware_concurrency();
  return h ? (int)h : 1;
}
static void pool_start(smash_ctx *c) {
#if !defined(SMASH_CUDA_SHIM)
  if (c->transport == 0) sched_calibrate_dma(c, c->slot[0].st_out);
  if (!c->pool.th.empty()) return;
  int n = c->host_threads;
  if (n <= 0) { if (const char *e = getenv("SMASH_HOST_THREADS")) n = atoi(e); }
  if (n <= 0) { n = host_cpu_count(); if (n > 16) n = 16; }
  for (int i = 0; i < n; ++i) c->pool.th.emplace_back(pool_main, &c->pool);
#endif
}

// ---- which way a read range's lines travel (SMASH transport, include/smash_b200.h)
static void sched_note_task(smash_ctx *c, double ns_per_rec) {
  std::lock_guard<std::mutex> lk(c->sched_mu);
  c->cpu_ns_per_rec += 0.05 * (ns_per_rec - c->cpu_ns_per_rec);      // under memory contention tasks stretch and the rate follows
}
static void sched_calibrate_dma(smash_ctx *c, cudaStream_t st) {
  if (c->dma_calibrated) return;
  c->dma_calibrated = true;
  const size_t n = 32u << 20;
  void *d = nullptr, *h = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaMalloc(&d, n) == cudaSuccess && cudaHostAlloc(&h, n, cudaHostAllocDefault) == cudaSuccess &&
      cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
    cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st);
    cudaEventRecord(e0, st);
    cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st);
    cudaEventRecord(e1, st);
    float ms = 0;
    if (cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0)
      c->dma_bytes_per_ms = 0.9 * (double)n / ms;             // (uploads share the link's host side)
  }
  cudaGetLastError();
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (h) cudaFreeHost(h);
  if (d) cudaFree(d);
  if (g_dbg) fprintf(stderr, "[smash-dbg] download rate %.1f GB/s\n", c->dma_bytes_per_ms / 1e6);
}
// true: compact transport for this range.  Both engines are modelled as queues: the download stream moves
// dma_bytes_per_ms, the pool builds a record in cpu_ns_per_rec per thread.  A compact range keeps the pool busy and
// takes 60 % less of the download stream, which is what later ranges queue behind: it is chosen as long as the pool
// would finish it no later than one full-text range after the moment the stream alone would have -- the two queues end
// up the same length, each engine working at its own measured rate (a host short of cores or of memory bandwidth
// stretches cpu_ns_per_rec and the ranges go back to the stream).
static bool sched_choose_compact(smash_ctx *c, uint64_t sam_bytes, uint64_t cmp_bytes, uint64_t recs, int ch) {
  if (c->transport == 2) return true;
  if (c->transport == 3) return (ch & 1) != 0;                 // tests: both ways inside one batch
  std::lock_guard<std::mutex> lk(c->sched_mu);
  const double now = now_ms();
  const double dma0 = c->dma_free_at > now ? c->dma_free_at : now, cpu0 = c->cpu_free_at > now ? c->cpu_free_at : now;
  const double full_cost = (double)sam_bytes / c->dma_bytes_per_ms;
  const double t_full = dma0 + full_cost;
  const double dma_c = dma0 + (double)(cmp_bytes + recs * sizeof(CmpMeta)) / c->dma_bytes_per_ms;
  const double threads = (double)(c->pool.th.size() + 1);
  const double t_cmp = (dma_c > cpu0 ? dma_c : cpu0) + (double)recs * c->cpu_ns_per_rec / 1e6 / threads;
  if (t_cmp <= t_full + full_cost) { c->dma_free_at = dma_c; c->cpu_free_at = t_cmp; return true; }
  c->dma_free_at = t_full;
  return false;
}
constexpr uint64_t EXPAND_GRAIN = 4096;                      // records per task
static void expand_dispatch(smash_ctx *c, Slot &s, int ch) {
  const uint64_t recs = s.cmp_recs[ch];
  if (!recs) return;
  std::vector<ExpandTask> ts;
  for (uint64_t f0 = 0; f0 < recs; f0 += EXPAND_GRAIN) ts.push_back(ExpandTask{&s, ch, f0, f0 + EXPAND_GRAIN < recs ? f0 + EXPAND_GRAIN : recs});
  s.pending.fetch_add((int)ts.size());
  { std::lock_guard<std::mutex> lk(c->pool.mu); c->pool.q.insert(c->pool.q.end(), ts.begin(), ts.end()); }
  c->pool.cv.notify_all();
}
// read ranges whose download has landed become expansion tasks; block: wait for the downloads still in flight
static int expand_poll(smash_ctx *c, Slot &s, bool block) {
  while (s.n_dispatched < s.n_ranges) {
    const int ch = s.n_dispatched;
    if (block) CU(cudaEventSynchronize(s.ev_d2h[ch]));
    else {
      const cudaError_t e = cudaEventQuery(s.ev_d2h[ch]);
      if (e == cudaErrorNotReady) return 0;
      if (e != cudaSuccess) return fail(SMASH_ERR_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
    }
    expand_dispatch(c, s, ch);
    ++s.n_dispatched;
  }
  return 0;
}
// every dispatched task has finished (the calling thread works too)
static void expand_drain(smash_ctx *c, Slot &s) {
  while (s.pending.load() > 0) {
    if (pool_try_run_one(c->pool)) continue;
    std::unique_lock<std::mutex> lk(s.mu);
    s.cv.wait_for(lk, std::chrono::microseconds(200), [&] { return s.pending.load() <= 0; });
  }
}

// The tail's appends must happen in the order the batches were submitted (first-wins dedupe, positions order), whichever
// slot worker gets there first: a batch takes the turn before its first append and passes it on when it has enqueued
// everything (or has nothing to append).
static void slot_begin_job(smash_ctx *c, Slot &s) {
  std::lock_guard<std::mutex> lk(c->turn_mu);
  s.job_seq = c->next_seq++; s.turn_held = false; s.turn_done = false;
}
static void tail_turn_acquire(smash_ctx *c, Slot &s) {
  if (s.turn_held || s.turn_done) return;
  std::unique_lock<std::mutex> lk(c->turn_mu);
  c->turn_cv.wait(lk, [&] { return c->tail_turn == s.job_seq; });
  s.turn_held = true;
}
static bool tail_turn_try(smash_ctx *c, Slot &s) {            // the turn, if it is this job's already (never waits)
  if (s.turn_held) return true;
  if (s.turn_done) return false;
  std::lock_guard<std::mutex> lk(c->turn_mu);
  if (c->tail_turn != s.job_seq) return false;
  s.turn_held = true;
  return true;
}
static void tail_turn_release(smash_ctx *c, Slot &s) {
  if (s.turn_done) return;
  std::unique_lock<std::mutex> lk(c->turn_mu);
  if (!s.turn_held) c->turn_cv.wait(lk, [&] { return c->tail_turn == s.job_seq; });
  c->tail_turn = s.job_seq + 1; s.turn_held = false; s.turn_done = true;
  lk.unlock();
  c->turn_cv.notify_all();
}
static void host_threads_stop(smash_ctx *c) {
  for (int i = 0; i < SMASH_N_SLOTS; ++i) {
    Slot &s = c->slot[i];
    if (s.worker.joinable()) {
      { std::lock_guard<std::mutex> lk(s.mu); s.quit = true; }
      s.cv.notify_all();
      s.worker.join();
    }
  }
  { std::lock_guard<std::mutex> lk(c->pool.mu); c->pool.quit = true; }
  c->pool.cv.notify_all();
  for (auto &t : c->pool.th) t.join();
  c->pool.th.clear();
}

static void slot_release(Slot &s) {
  s.names.release(); s.seq.release(); s.qual.release(); s.opt.release(); s.name_off.release();
  s.seq_off.release(); s.opt_off.release(); s.read_flag.release(); s.match_slots.release(); s.mem_stage.release();
  s.match_cnt.release(); s.item_slots.release(); s.rec_slots.release(); s.sums.release();
  s.nrec.release(); s.rec_base.release(); s.rec_read.release(); s.rec_bytes.release(); s.rec_off.release(); s.sam_total.release(); s.blk_sums2.release(); s.blk_sums.release(); s.sam.release(); s.flags.release();
  s.csr_off.release(); s.csr_triples.release(); s.long_scratch.release(); s.slot_off.release(); s.aln_scr.release(); s.ord_scr.release(); s.tmp32.release(); s.h_sam.release(); s.h_csr_off.release();
  s.h_matches.release(); s.h_small.release(); s.h_flag.release(); s.surv.release(); s.surv_cnt.release(); s.lc.release(); s.slow.release();
  s.sort_abs.release(); s.sort_off.release(); s.sort_flag.release(); s.sort_tmp.release(); s.sort_perm.release(); s.sort_bytes.release();
  for (int f = 0; f < 2; ++f) { s.ing_raw[f].release(); s.ing_ls[f].release(); s.ing_hdr[f].release(); }
  s.ing_hdr_flag.release(); s.ing_blk64.release(); s.ing_blk32.release(); s.ing_blk4.release(); s.ing_pre.release();
  s.ing_recs.release(); s.ing_scal.release(); s.h_ing.release();
  if (s.ev_ing0) cudaEventDestroy(s.ev_ing0);
  if (s.ev_ing1) cudaEventDestroy(s.ev_ing1);
  if (s.ev0) cudaEventDestroy(s.ev0);
  if (s.ev1) cudaEventDestroy(s.ev1);
  for (int e = 0; e < N_EVS; ++e) if (s.evs[e]) cudaEventDestroy(s.evs[e]);
  for (int e = 0; e < MAX_CHUNKS; ++e) {
    if (s.ev_in[e]) cudaEventDestroy(s.ev_in[e]);
    if (s.ev_emit[e]) cudaEventDestroy(s.ev_emit[e]);
    if (s.ev_d2h[e]) cudaEventDestroy(s.ev_d2h[e]);
    s.cmp[e].release(); s.cmeta[e].release(); s.h_cmp[e].release(); s.h_cmeta[e].release();
  }
  s.cmp_bytes.release(); s.cmp_off.release();
  if (s.ev_out) cudaEventDestroy(s.ev_out);
  if (s.st_in) cudaStreamDestroy(s.st_in);
  if (s.st_tail) cudaStreamDestroy(s.st_tail);
  if (s.ev_tail) cudaEventDestroy(s.ev_tail);
  if (s.st_out && s.owns_out) cudaStreamDestroy(s.st_out);
  if (s.st) cudaStreamDestroy(s.st);
}

extern "C" void smash_ctx_destroy(smash_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  host_threads_stop(c);
  smash_comm_destroy(c);
  for (int s = 0; s < SMASH_N_SLOTS; ++s) slot_release(c->slot[s]);
  tail_release(&c->tail);
  void *ptrs[] = {c->seed_irr, c->descr8, c->ext, c->text_alloc, c->sa, c->isa, c->lcp, c->lcp_m, c->uniq, c->seed, c->startpos, c->sizes,
                  c->descr, c->descr_off, c->alpha, c->mapbody, c->chrom_off32, c->chrom_abs64};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (c->own_index) delete c->own_index;
  delete c;
}

extern "C" int smash_ctx_load_mappability(smash_ctx *c, const uint8_t *body, uint64_t n) {
  if (!c || !body) return fail(SMASH_ERR_ARG, "null argument");
  // map.bin exists only for -rcref indexes (mummer.cpp:145); the L/R tags and the tail index it by forward chromosome
  if (!c->hix->rcref) return fail(SMASH_ERR_ARG, "mappability requires an -rcref index");
  CU(cudaSetDevice(c->device));
  if (c->mapbody) { cudaFree(c->mapbody); c->mapbody = nullptr; }
  int rc = dmalloc((void **)&c->mapbody, n, &c->index_bytes);
  if (rc) return rc;
  CU(cudaMemcpy(c->mapbody, body, n, cudaMemcpyHostToDevice));
  c->dix.mapbody = c->mapbody; c->dix.map_bytes = n;
  return 0;
}

extern "C" int smash_ctx_build_mappability(smash_ctx *c, uint8_t *body, uint64_t cap) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  if (!c->hix->isa && !c->isa) return fail(SMASH_ERR_STATE, "mappability needs the .isa.bin array");
  if (!c->hix->rcref) return fail(SMASH_ERR_ARG, "-mappability requires -rcref");   // mummer.cpp:145
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->slot[0].st;
  uint64_t total = 0;
  for (size_t i = 0; i < c->hix->sizes.size(); i += 2) total += 2 * c->hix->sizes[i];
  bool temp_isa = false;
  if (!c->isa) {
    int rc = dmalloc(&c->isa, c->dix.N * c->dix.w, nullptr); if (rc) return rc;
    CU(cudaMemcpy(c->isa, c->hix->isa, c->dix.N * c->dix.w, cudaMemcpyHostToDevice));
    c->dix.isa = c->isa; temp_isa = true;
  }
  if (c->mapbody) { cudaFree(c->mapbody); c->mapbody = nullptr; }
  int rc = dmalloc((void **)&c->mapbody, total, &c->index_bytes);
  if (rc) return rc;
  c->launches += launch_mappability(c->dix, nullptr, c->mapbody, st);
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  if (temp_isa && c->prm.mode != SMASH_MODE_MEM) { cudaFree(c->isa); c->isa = nullptr; c->dix.isa = nullptr; }
  c->dix.mapbody = c->mapbody; c->dix.map_bytes = total;
  if (body) {
    if (cap < total) return fail(SMASH_ERR_ARG, "mappability buffer too small: need %llu bytes", (unsigned long long)total);
    CU(cudaMemcpy(body, c->mapbody, total, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ------------------------------------------------------------------ batches

// Device buffers of a slot for a batch of n reads with blobs of the given sizes, and the read ranges of the
// chunked pipeline.  The batch itself arrives either by H2D copies (slot_prepare) or from the device-side
// input stage (ingest_text).
static int slot_reserve(smash_ctx *c, Slot &s, uint64_t n, size_t name_bytes, size_t seq_bytes, size_t opt_bytes, int chunks) {
  double t_prep = now_ms();
  int rc;
  if ((rc = s.names.ensure(name_bytes + 16)) || (rc = s.seq.ensure(seq_bytes + 16)) || (rc = s.qual.ensure(seq_bytes + 16)) ||
      (rc = s.name_off.ensure(n + 1)) || (rc = s.seq_off.ensure(n + 1)) || (rc = s.read_flag.ensure(n + 1)))
    return rc;
  if (opt_bytes && ((rc = s.opt.ensure(opt_bytes + 16)) || (rc = s.opt_off.ensure(n + 1)))) return rc;
  if (s.cap == 0) s.cap = 24;
  // split search (k_mam_search parks candidates, k_mam_verify extends them): pays off on large references,
  // where the 4+4 filter exists and every read has a few dozen seed hits; small references verify in place
  s.split = c->prm.mode != SMASH_MODE_MEM && !g_no_split && (c->dix.ext != nullptr || g_force_split);
  s.seed_ok = s.split && c->dix.ext != nullptr && c->sp.fast_ok;
  if (s.split && ((rc = s.surv.ensure(n * SURV_CAP + 1)) || (rc = s.surv_cnt.ensure(n + 1)) || (rc = s.lc.ensure(seq_bytes + 32 * n + 64)) ||
                  (rc = s.slow.ensure(n + 1))))
    return rc;
  if ((rc = s.match_slots.ensure(n * s.cap)) || (rc = s.match_cnt.ensure(n + 1)) || (rc = s.item_slots.ensure(n * s.cap)) ||
      (rc = s.rec_slots.ensure(n * s.cap)) || (rc = s.sums.ensure(n + 2)) || (rc = s.nrec.ensure(n + 1)) ||
      (rc = s.rec_base.ensure(n + 2)) || (rc = s.rec_read.ensure(n * s.cap + 1)) || (rc = s.rec_bytes.ensure(n * s.cap + 1)) ||
      (rc = s.rec_off.ensure(n * s.cap + 2)) || (rc = s.sam_total.ensure(2)) || (rc = s.blk_sums.ensure(n / 2048 + 8)) ||
      (rc = s.blk_sums2.ensure(n * s.cap / 2048 + 8)) || (rc = s.flags.ensure(N_FLAGS)) ||
      (rc = s.h_small.ensure(32)))
    return rc;
  if (s.compact && ((rc = s.cmp_bytes.ensure(n * s.cap + 1)) || (rc = s.cmp_off.ensure(n * s.cap + 2)))) return rc;
  DBG_T("  prepare:ensure", t_prep);
  // read ranges of the chunked pipeline (even boundaries: mates stay together)
  {
    const uint64_t by_size = c->chunk_min_reads ? n / c->chunk_min_reads : n;      // ranges of at least chunk_min_reads reads
    const int lim = chunks > MAX_CHUNKS ? MAX_CHUNKS : chunks;
    s.n_chunks = (chunks > 1 && by_size >= 1) ? (int)(by_size < (uint64_t)lim ? (by_size > 1 ? by_size : 1) : (uint64_t)lim) : 1;
  }
  {
    const uint64_t per = ((n + s.n_chunks - 1) / s.n_chunks + 1) & ~1ull;
    for (int ch = 0; ch <= s.n_chunks; ++ch) { const uint64_t r = per * ch; s.chunk_r[ch] = r < n ? r : n; }
    s.chunk_r[s.n_chunks] = n;
  }
  s.name_bytes = name_bytes; s.seq_bytes = seq_bytes; s.opt_bytes = opt_bytes;
  for (int ch = 0; ch < MAX_CHUNKS; ++ch) s.chunk_name_bytes[ch] = name_bytes;
  s.bd.n_reads = n; s.bd.names = s.names.p; s.bd.name_off = s.name_off.p; s.bd.seq = s.seq.p; s.bd.qual = s.qual.p;
  s.bd.seq_off = s.seq_off.p; s.bd.opt = opt_bytes ? s.opt.p : nullptr; s.bd.opt_off = opt_bytes ? s.opt_off.p : nullptr;
  s.bd.read_flag = s.read_flag.p;
  s.n_reads = n;
  s.long_q = 0;               // reads longer than MAXQ_FAST: the search kernel reports them, run_range re-runs with scratch
  return 0;
}

static int slot_prepare(smash_ctx *c, Slot &s, const smash_batch *b, bool copy, int chunks = 1) {
  const uint64_t n = b->n_reads;
  const size_t name_bytes = n ? (size_t)b->name_off[n] : 0, seq_bytes = n ? (size_t)b->seq_off[n] : 0;
  const size_t opt_bytes = (b->opt && n) ? (size_t)b->opt_off[n] : 0;
  int rc;
  if ((rc = slot_reserve(c, s, n, name_bytes, seq_bytes, opt_bytes, chunks))) return rc;
  double t_prep = now_ms();
  if (copy && n) {
    cudaStream_t in = s.n_chunks > 1 ? s.st_in : s.st;
    // read_flag is usually computed on the fly by the caller (2 B/read): if it sits in pageable memory its
    // copy would block this thread behind every upload queued before it, so it goes through pinned staging
    const uint16_t *flag_src = b->read_flag;
    {
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, b->read_flag) != cudaSuccess || at.type == cudaMemoryTypeUnregistered) {
        cudaGetLastError();
        if ((rc = s.h_flag.ensure(n + 1))) return rc;
        memcpy(s.h_flag.p, b->read_flag, 2 * n);
        flag_src = s.h_flag.p;
      }
    }
    for (int ch = 0; ch < s.n_chunks; ++ch) {
      const uint64_t r0 = s.chunk_r[ch], r1 = s.chunk_r[ch + 1];
      if (r1 > r0) {
        const size_t n0 = (size_t)b->name_off[r0], n1 = (size_t)b->name_off[r1], q0 = (size_t)b->seq_off[r0], q1 = (size_t)b->seq_off[r1];
        s.chunk_name_bytes[ch] = n1 - n0;
        CU(cudaMemcpyAsync(s.names.p + n0, b->names + n0, n1 - n0, cudaMemcpyHostToDevice, in));
        CU(cudaMemcpyAsync(s.seq.p + q0, b->seq + q0, q1 - q0, cudaMemcpyHostToDevice, in));
        CU(cudaMemcpyAsync(s.qual.p + q0, b->qual + q0, q1 - q0, cudaMemcpyHostToDevice, in));
        CU(cudaMemcpyAsync(s.name_off.p + r0, b->name_off + r0, 8 * (r1 - r0 + 1), cudaMemcpyHostToDevice, in));
        CU(cudaMemcpyAsync(s.seq_off.p + r0, b->seq_off + r0, 8 * (r1 - r0 + 1), cudaMemcpyHostToDevice, in));
        CU(cudaMemcpyAsync(s.read_flag.p + r0, flag_src + r0, 2 * (r1 - r0), cudaMemcpyHostToDevice, in));
        s.io_h2d += (n1 - n0) + 2 * (q1 - q0) + 16 * (r1 - r0 + 1) + 2 * (r1 - r0);
        if (opt_bytes) {
          const size_t o0 = (size_t)b->opt_off[r0], o1 = (size_t)b->opt_off[r1];
          CU(cudaMemcpyAsync(s.opt.p + o0, b->opt + o0, o1 - o0, cudaMemcpyHostToDevice, in));
          CU(cudaMemcpyAsync(s.opt_off.p + r0, b->opt_off + r0, 8 * (r1 - r0 + 1), cudaMemcpyHostToDevice, in));
          s.io_h2d += (o1 - o0) + 8 * (r1 - r0 + 1);
        }
      }
      if (s.n_chunks > 1) CU(cudaEventRecord(s.ev_in[ch], in));
    }
  }
  DBG_T("  prepare:copies enqueued", t_prep);
  s.first_pair = b->first_pair_ordinal;
  return 0;
}

static WorkDev work_of(Slot &s) {
  WorkDev w{};
  w.cap = s.cap; w.slot_off = s.csr ? s.slot_off.p : nullptr; w.slots_total = s.csr ? s.slots_total : s.n_reads * (uint64_t)s.cap;
  w.aln_scratch = s.aln_scr.p; w.ord_scratch = s.ord_scr.p;
  if (!s.csr && s.surv.p && s.split) {
    w.surv = s.surv.p; w.surv_cnt = s.surv_cnt.p; w.lc = s.lc.p;
    // the lean seed stage needs the ext table and a seed rare enough for the anchor path; SMASH_NO_SEED_KERNEL: A/B switch
    static const bool no_seed = getenv("SMASH_NO_SEED_KERNEL") != nullptr;
    if (!no_seed && s.slow.p && s.seed_ok) w.slow = s.slow.p;
  }
  w.long_scratch = s.long_q ? s.long_scratch.p : nullptr; w.long_q = s.long_q; w.match_slots = s.match_slots.p; w.match_cnt = s.match_cnt.p; w.item_slots = s.item_slots.p;
  w.rec_slots = s.rec_slots.p; w.sums = s.sums.p; w.nrec = s.nrec.p; w.rec_base = s.rec_base.p; w.rec_read = s.rec_read.p; w.rec_bytes = s.rec_bytes.p; w.rec_off = s.rec_off.p; w.sam_total = s.sam_total.p; w.blk_sums2 = s.blk_sums2.p;
  w.sort_abs = s.sort_abs.p; w.sort_flag = s.sort_flag.p; w.sort_perm = s.sort_perm.p; w.sort_bytes = s.sort_bytes.p; w.sort_off = s.sort_off.p;
  w.sort_tmp = s.sort_tmp.p; w.sort_tmp_bytes = s.sort_tmp.cap;
  if (s.compact) { w.cmp_bytes = s.cmp_bytes.p; w.cmp_off = s.cmp_off.p; w.sam_base = s.sam_base; }
  w.blk_sums = s.blk_sums.p; w.sam = s.sam.p ? s.sam.p + s.sam_base : nullptr; w.sam_cap = s.sam.cap > s.sam_base ? s.sam.cap - s.sam_base : 0; w.flags = s.flags.p;
  return w;
}

#define MARK(stage) do { if (s.n_evs < N_EVS) { CU(cudaEventRecord(s.evs[s.n_evs], s.st)); s.ev_stage[s.n_evs++] = (stage); } } while (0)

// One read range of the slot's batch (s.bd / s.n_reads are the range's view):
// search -> records -> sizes/scan -> (sync for the byte total) -> emit [-> D2H] [-> tail append].
// The only host synchronisation inside is the read of the published SAM size.
static int run_range(smash_ctx *c, Slot &s, int want, bool to_host, int ch, uint64_t n_full) {
  const uint64_t n = s.n_reads;
  const bool chunked = s.n_chunks > 1;
  const int evs_at_entry = s.n_evs;
  bool exact_csr = false;
  for (int attempt = 0;; ++attempt) {
    WorkDev w = work_of(s);
    s.n_evs = evs_at_entry;
    CU(cudaMemsetAsync(s.flags.p, 0, sizeof(uint32_t) * N_FLAGS, s.st));
    if (ch == 0 && !chunked) CU(cudaEventRecord(s.ev0, s.st));
    else MARK(-1);                                           // stage timers restart here (after the upload wait)
    int nl;
    if (c->prm.mode == SMASH_MODE_MEM) {
      // K2: count pass -> slot offsets (one spare slot per read) -> exact-size buffers -> write pass
      int rc;
      if ((rc = s.slot_off.ensure(n + 2)) || (rc = s.tmp32.ensure(n + 1)) || (rc = s.mem_stage.ensure(n * MEM_STAGE + 1))) return rc;
      nl = launch_mem_count(c->dix, s.bd, c->sp, c->prm.min_len, s.match_cnt.p, s.mem_stage.p, s.st);
      nl += launch_slot_offsets(s.match_cnt.p, n, s.tmp32.p, s.blk_sums.p, s.slot_off.p, s.st);
      CU(cudaMemcpyAsync(s.h_small.p + 16, s.slot_off.p + n, 8, cudaMemcpyDeviceToHost, s.st));
      CU(cudaStreamSynchronize(s.st));
      const uint64_t total = s.h_small.p[16];
      s.slots_total = total; s.csr = true;
      if ((rc = s.match_slots.ensure(total + 1)) || (rc = s.item_slots.ensure(total + 1)) || (rc = s.rec_slots.ensure(total + 1)) ||
          (rc = s.aln_scr.ensure(total + 1)) || (rc = s.ord_scr.ensure(total + 1)) || (rc = s.rec_read.ensure(total + 1)) ||
          (rc = s.rec_bytes.ensure(total + 1)) || (rc = s.rec_off.ensure(total + 2)) || (rc = s.blk_sums2.ensure(total / 2048 + 8)))
        return rc;
      if (s.compact && ((rc = s.cmp_bytes.ensure(total + 1)) || (rc = s.cmp_off.ensure(total + 2)))) return rc;
      w = work_of(s);
      nl += launch_mem_write(c->dix, s.bd, c->sp, c->prm.min_len, s.slot_off.p, s.match_slots.p, s.mem_stage.p, s.match_cnt.p, s.st);
    } else if (exact_csr) {
      // K1c: a read of this range has more matches than the anchor kernels stage (STAGE_CAP): exact per-start search with
      // CSR slots, any number of matches per read (count pass -> slot offsets -> exact-size buffers -> write pass)
      int rc;
      if (s.long_q < MAXQ_FAST) s.long_q = MAXQ_FAST;
      if ((rc = s.long_scratch.ensure((size_t)148 * 8 * 8 * (size_t)(s.long_q + P_FRONT + P_BACK + 8))) || (rc = s.slot_off.ensure(n + 2)) ||
          (rc = s.tmp32.ensure(n + 1)))
        return rc;
      s.csr = false;
      w = work_of(s);
      nl = launch_mam_exact(c->dix, s.bd, w, c->sp, s.match_cnt.p, s.st);
      nl += launch_slot_offsets(s.match_cnt.p, n, s.tmp32.p, s.blk_sums.p, s.slot_off.p, s.st);
      CU(cudaMemcpyAsync(s.h_small.p + 16, s.slot_off.p + n, 8, cudaMemcpyDeviceToHost, s.st));
      CU(cudaStreamSynchronize(s.st));
      const uint64_t total = s.h_small.p[16];
      s.slots_total = total; s.csr = true;
      if ((rc = s.match_slots.ensure(total + 1)) || (rc = s.item_slots.ensure(total + 1)) || (rc = s.rec_slots.ensure(total + 1)) ||
          (rc = s.aln_scr.ensure(total + 1)) || (rc = s.ord_scr.ensure(total + 1)) || (rc = s.rec_read.ensure(total + 1)) ||
          (rc = s.rec_bytes.ensure(total + 1)) || (rc = s.rec_off.ensure(total + 2)) || (rc = s.blk_sums2.ensure(total / 2048 + 8)))
        return rc;
      if (s.compact && ((rc = s.cmp_bytes.ensure(total + 1)) || (rc = s.cmp_off.ensure(total + 2)))) return rc;
      w = work_of(s);
      nl += launch_mam_exact(c->dix, s.bd, w, c->sp, nullptr, s.st);
    } else {
      if (s.csr) { s.csr = false; w = work_of(s); }           // an earlier range of this slot ran on CSR slots (K1c)
      nl = launch_mam_search(c->dix, s.bd, w, c->sp, s.st);
    }
    s.launches += nl;
    MARK(0);
    if (!s.csr) {
      const int nv = launch_mam_verify(c->dix, s.bd, w, c->sp, s.st);
      s.launches += nv;
      if (nv) MARK(7);
    }
    s.launches += launch_records(c->dix, s.bd, w, c->sp, s.st);
    MARK(1);
    s.launches += launch_sizes_scan(c->dix, s.bd, w, c->sp, s.st);
    MARK(2);
    s.launches += launch_publish(s.h_small.p, (const uint64_t *)s.sam_total.p, (const uint64_t *)(s.rec_base.p + n), s.flags.p, s.st);
    { const double ts = now_ms(); CU(cudaStreamSynchronize(s.st)); DBG_T("  run:sync for sizes", ts); }
    const uint32_t *fl = (const uint32_t *)(s.h_small.p + 1);
    if (fl[FLAG_LONGQ] > (uint32_t)s.long_q && c->prm.mode != SMASH_MODE_MEM) {
      // reads longer than the shared-memory staging buffer get per-warp scratch in HBM (exact search path)
      const uint32_t max_q = fl[FLAG_LONGQ];
      if (max_q > 60000 || attempt > 3) return fail(SMASH_ERR_ARG, "read of %u bases: reads longer than 60000 are not supported", max_q);
      int rc;
      s.long_q = (int)max_q;
      if ((rc = s.long_scratch.ensure((size_t)148 * 8 * 8 * (size_t)(max_q + P_FRONT + P_BACK + 8)))) return rc;
      continue;                                              // rerun the range with the long-read kernel
    }
    if (fl[FLAG_LONGREAD]) return fail(SMASH_ERR_STATE, "%u long reads could not be staged", fl[FLAG_LONGREAD]);
    if (fl[FLAG_OVERFLOW]) {
      const uint32_t need = fl[FLAG_MAXCNT];
      if (exact_csr || attempt > 4) return fail(SMASH_ERR_DATA, "a read produced more than 65535 matches (%u)", need);
      if (need > (uint32_t)STAGE_CAP) { exact_csr = true; continue; }   // rerun the range on the exact CSR path (K1c)
      s.cap = (int)need + 8 > STAGE_CAP ? STAGE_CAP : (int)need + 8;
      int rc;
      if ((rc = s.match_slots.ensure(n_full * s.cap)) || (rc = s.item_slots.ensure(n_full * s.cap)) || (rc = s.rec_slots.ensure(n_full * s.cap)) ||
          (rc = s.rec_read.ensure(n_full * s.cap + 1)) || (rc = s.rec_bytes.ensure(n_full * s.cap + 1)) || (rc = s.rec_off.ensure(n_full * s.cap + 2)) ||
          (rc = s.blk_sums2.ensure(n_full * s.cap / 2048 + 8))) return rc;
      if (s.compact && ((rc = s.cmp_bytes.ensure(n_full * s.cap + 1)) || (rc = s.cmp_off.ensure(n_full * s.cap + 2)))) return rc;
      continue;                                              // rerun the range with wider slots
    }
    if (fl[FLAG_MAPERR] && ((want & SMASH_WANT_TAIL) || c->prm.tag_mappability))
      return fail(SMASH_ERR_DATA, "left/right mappability too big for %u records (mappability_tag.cpp:107-113 throws here)", fl[FLAG_MAPERR]);
    break;
  }
  const uint64_t bytes = s.h_small.p[0], recs = s.h_small.p[8];
  // The tail append reads the records only, so it can run NEXT TO the emit kernels (small latency-bound kernels beside
  // issue-bound ones) on the slot's side stream: everything it reads was complete at the synchronisation above.  Not with
  // SMASH_WANT_SORTED (the sort re-orders record arrays), and only if the append turn is this batch's already (a worker
  // never delays its emit launches waiting for an earlier batch).
  bool tail_early = false;
  if ((want & SMASH_WANT_TAIL) && !g_no_tail_overlap && !(want & SMASH_WANT_SORTED) && tail_turn_try(c, s)) {
    int rc = tail_accumulate(&c->tail, c->dix, s.bd, work_of(s), recs, s.chunk_name_bytes[ch], s.st_tail, &s.launches);
    if (rc) return fail(rc, "tail: %s", tail_error());
    CU(cudaEventRecord(s.ev_tail, s.st_tail));
    tail_early = true;
  }
  const bool range_compact = (want & SMASH_WANT_SAM) && s.compact && sched_choose_compact(c, bytes, s.h_small.p[9], recs, ch);
  if (g_dbg && s.compact) fprintf(stderr, "[smash-dbg]   range %d of slot %d: %s\n", ch, (int)(&s - c->slot), range_compact ? "compact" : "full text");
  if (range_compact) {
    // compact transport: only the text the GPU computes + one CmpMeta per record cross PCIe (compact.h)
    int rc;
    const uint64_t cbytes = s.h_small.p[9];
    if (cbytes >= 0xffffff00ull) return fail(SMASH_ERR_ARG, "read range too large for the compact transport (%llu bytes)", (unsigned long long)cbytes);
    if ((rc = s.cmp[ch].ensure(cbytes + 64)) || (rc = s.cmeta[ch].ensure(recs + 1)) || (rc = s.h_cmp[ch].ensure(cbytes + 64)) ||
        (rc = s.h_cmeta[ch].ensure(recs + 1)))
      return rc;
    const uint64_t need = s.sam_base + bytes + 64;
    const uint64_t guess = (chunked && ch == 0 && n) ? (uint64_t)((double)bytes * ((double)n_full / (double)n) * 1.03) + 4096 : 0;
    if (need > s.h_sam.cap) {
      expand_drain(c, s);                                     // lines of earlier ranges are being written into the old buffer
      if (s.sam_base) CU(cudaStreamSynchronize(s.st_out));
      if ((rc = s.h_sam.grow_keep(need > guess ? need : guess, s.sam_base))) return rc;
    }
    if ((want & SMASH_WANT_SORTED) && recs) {
      if (recs >= 0xffffffffull) return fail(SMASH_ERR_ARG, "too many records in one batch for SMASH_WANT_SORTED");
      size_t tmp_bytes = 0;
      launch_record_sort(c->dix, s.bd, work_of(s), recs, s.st, &tmp_bytes);
      if ((rc = s.sort_abs.ensure(recs + 1)) || (rc = s.sort_off.ensure(recs + 2)) || (rc = s.sort_flag.ensure(recs + 1)) ||
          (rc = s.sort_perm.ensure(recs + 1)) || (rc = s.sort_bytes.ensure(recs + 1)) || (rc = s.sort_tmp.ensure(tmp_bytes + 16)) ||
          (rc = s.blk_sums2.ensure(recs / 2048 + 8)))
        return rc;
      s.launches += launch_record_sort(c->dix, s.bd, work_of(s), recs, s.st, nullptr);
      s.sorted = true;
    }
    WorkDev w = work_of(s);
    w.cmp = s.cmp[ch].p; w.cmeta = s.cmeta[ch].p;
    s.launches += launch_emit_compact(c->dix, s.bd, w, c->sp, s.st, recs);
    MARK(3);
    if (s.sorted) s.launches += launch_publish(s.h_small.p, (const uint64_t *)s.sam_total.p, (const uint64_t *)(s.rec_base.p + n), s.flags.p, s.st);
    cudaStream_t out = s.st;
    if (chunked) {
      CU(cudaEventRecord(s.ev_emit[ch], s.st));
      CU(cudaStreamWaitEvent(s.st_out, s.ev_emit[ch], 0));
      out = s.st_out;
    }
    if (g_dbg && s.tl[1] && ch == 0) cudaEventRecord(s.tl[1], out);
    if (recs) {
      CU(cudaMemcpyAsync(s.h_cmeta[ch].p, s.cmeta[ch].p, recs * sizeof(CmpMeta), cudaMemcpyDeviceToHost, out));
      CU(cudaMemcpyAsync(s.h_cmp[ch].p, s.cmp[ch].p, cbytes, cudaMemcpyDeviceToHost, out));
    }
    if (g_dbg && s.tl[2] && ch == s.n_chunks - 1) cudaEventRecord(s.tl[2], out);
    CU(cudaEventRecord(s.ev_d2h[ch], out));
    s.cmp_recs[ch] = recs; s.n_ranges = ch + 1;
    s.io_d2h += recs * sizeof(CmpMeta) + cbytes;
  } else if (want & SMASH_WANT_SAM) {
    int rc;
    const uint64_t need = s.sam_base + bytes + 64;
    // first range of a chunked batch: size both buffers for the whole batch from this range's density
    const uint64_t guess = (chunked && ch == 0 && n) ? (uint64_t)((double)bytes * ((double)n_full / (double)n) * 1.03) + 4096 : 0;
    if (need > s.sam.cap) {
      if (s.sam_base) CU(cudaStreamSynchronize(s.st_out));   // earlier ranges are on the host already; their device copy may go
      if ((rc = s.sam.ensure(need > guess ? need : guess))) return rc;
    }
    if ((want & SMASH_WANT_SORTED) && recs) {
      // K5: the batch is one chunk of the reference's OutputSorter -> its lines in MemSam::operator< order
      if (recs >= 0xffffffffull) return fail(SMASH_ERR_ARG, "too many records in one batch for SMASH_WANT_SORTED");
      size_t tmp_bytes = 0;
      launch_record_sort(c->dix, s.bd, work_of(s), recs, s.st, &tmp_bytes);
      if ((rc = s.sort_abs.ensure(recs + 1)) || (rc = s.sort_off.ensure(recs + 2)) || (rc = s.sort_flag.ensure(recs + 1)) ||
          (rc = s.sort_perm.ensure(recs + 1)) || (rc = s.sort_bytes.ensure(recs + 1)) || (rc = s.sort_tmp.ensure(tmp_bytes + 16)) ||
          (rc = s.blk_sums2.ensure(recs / 2048 + 8)))
        return rc;
      s.launches += launch_record_sort(c->dix, s.bd, work_of(s), recs, s.st, nullptr);
      s.sorted = true;
    }
    WorkDev w = work_of(s);
    s.launches += launch_emit_text(c->dix, s.bd, w, c->sp, s.st, recs);
    MARK(3);
    s.launches += launch_emit_copy(s.bd, w, s.st, recs);
    MARK(6);
    if (s.sorted) s.launches += launch_publish(s.h_small.p, (const uint64_t *)s.sam_total.p, (const uint64_t *)(s.rec_base.p + n), s.flags.p, s.st);
    if (to_host) {
      if (need > s.h_sam.cap) {
        if (s.sam_base) CU(cudaStreamSynchronize(s.st_out));
        if (s.compact) expand_drain(c, s);
        if ((rc = s.h_sam.grow_keep(need > guess ? need : guess, s.sam_base))) return rc;
      }
      cudaStream_t out = s.st;
      if (chunked) {
        CU(cudaEventRecord(s.ev_emit[ch], s.st));
        CU(cudaStreamWaitEvent(s.st_out, s.ev_emit[ch], 0));
        out = s.st_out;
      }
      if (g_dbg && s.tl[1] && ch == 0) cudaEventRecord(s.tl[1], out);
      CU(cudaMemcpyAsync(s.h_sam.p + s.sam_base, s.sam.p + s.sam_base, bytes, cudaMemcpyDeviceToHost, out));
      s.io_d2h += bytes;
      if (s.compact) { CU(cudaEventRecord(s.ev_d2h[ch], out)); s.cmp_recs[ch] = 0; s.n_ranges = ch + 1; }
      if (g_dbg && s.tl[2] && ch == s.n_chunks - 1) cudaEventRecord(s.tl[2], out);
    }
  }
  if (want & SMASH_WANT_MATCHES) {
    int rc;
    if ((rc = s.csr_off.ensure(n + 2))) return rc;
    const size_t mslots = s.csr ? (size_t)s.slots_total : n * (size_t)s.cap;
    if ((rc = s.csr_triples.ensure(3 * mslots + 8))) return rc;
    WorkDev w = work_of(s);
    s.launches += launch_match_csr(s.bd, w, s.csr_off.p, s.csr_triples.p, s.blk_sums.p, s.st);
    MARK(4);
    if ((rc = s.h_csr_off.ensure(n + 2)) || (rc = s.h_matches.ensure(mslots + 8))) return rc;
    CU(cudaMemcpyAsync(s.h_csr_off.p, s.csr_off.p, 8 * (n + 1), cudaMemcpyDeviceToHost, s.st));
    CU(cudaMemcpyAsync(s.h_matches.p, s.csr_triples.p, 24 * mslots, cudaMemcpyDeviceToHost, s.st));
    s.io_d2h += 8 * (n + 1) + 24 * mslots;
  }
  if (tail_early) {
    CU(cudaStreamWaitEvent(s.st, s.ev_tail, 0));              // the range is done when its append is
    MARK(5);
  } else if (want & SMASH_WANT_TAIL) {
    tail_turn_acquire(c, s);                                  // appends happen in the order the batches were submitted
    const double tt = now_ms();
    int rc = tail_accumulate(&c->tail, c->dix, s.bd, work_of(s), recs, s.chunk_name_bytes[ch], s.st, &s.launches);
    DBG_T("  run:tail_accumulate", tt);
    if (rc) return fail(rc, "tail: %s", tail_error());
    MARK(5);
  }
  s.sam_bytes += bytes; s.n_records += recs;
  if (want & SMASH_WANT_SAM) s.sam_base += bytes;
  return 0;
}

static int slot_run(smash_ctx *c, Slot &s, int want, bool to_host) {
  s.sam_bytes = 0; s.n_matches = 0; s.n_records = 0; s.want = want; s.sam_base = 0; s.n_evs = 0; s.sorted = false;
  s.n_ranges = 0; s.n_dispatched = 0;
  for (int ch = 0; ch < MAX_CHUNKS; ++ch) s.cmp_recs[ch] = 0;
  if (!s.n_reads) return 0;
  if (s.n_chunks > 1 && (!to_host || (want & SMASH_WANT_MATCHES) || c->prm.mode == SMASH_MODE_MEM))
    return fail(SMASH_ERR_STATE, "internal: chunked batch on a path that cannot take one");
  const BatchDev full = s.bd;
  const uint64_t n_full = s.n_reads;
  int rc = 0;
  if (s.n_chunks > 1) CU(cudaEventRecord(s.ev0, s.st));
  for (int ch = 0; ch < s.n_chunks && !rc; ++ch) {
    const uint64_t r0 = s.n_chunks > 1 ? s.chunk_r[ch] : 0, r1 = s.n_chunks > 1 ? s.chunk_r[ch + 1] : n_full;
    if (r1 <= r0) continue;
    s.bd = full; s.bd.n_reads = r1 - r0;
    s.bd.name_off = full.name_off + r0; s.bd.seq_off = full.seq_off + r0; s.bd.read_flag = full.read_flag + r0;
    if (full.opt_off) s.bd.opt_off = full.opt_off + r0;
    s.n_reads = r1 - r0;
    if (s.n_chunks > 1) { cudaError_t e = cudaStreamWaitEvent(s.st, s.ev_in[ch], 0); if (e != cudaSuccess) { rc = fail(SMASH_ERR_CUDA, "%s", cudaGetErrorString(e)); break; } }
    rc = run_range(c, s, want, to_host, ch, n_full);
    if (!rc && s.compact) rc = expand_poll(c, s, false);      // ranges already on the host: their lines are built meanwhile
  }
  s.bd = full; s.n_reads = n_full;
  if (rc) return rc;
  if (s.n_chunks > 1) {
    CU(cudaEventRecord(s.ev_out, s.st_out));
    CU(cudaStreamWaitEvent(s.st, s.ev_out, 0));
  }
  s.sam_base = 0;
  CU(cudaEventRecord(s.ev1, s.st));
  return 0;
}

// device time of the input stage; call after the slot's stream has been synchronised
static void ingest_collect(smash_ctx *c, Slot &s) {
  if (!s.ing_timed) return;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, s.ev_ing0, s.ev_ing1) == cudaSuccess) c->ingest_ms += ms; else cudaGetLastError();
  s.ing_timed = false;
}

static int slot_finish(smash_ctx *c, Slot &s, smash_result *res) {
  CU(cudaStreamSynchronize(s.st));
  CU(cudaGetLastError());
  c->launches += s.launches; c->io_h2d += s.io_h2d; c->io_d2h += s.io_d2h;
  s.launches = 0; s.io_h2d = 0; s.io_d2h = 0;
  ingest_collect(c, s);
  if (s.sorted && ((const uint32_t *)(s.h_small.p + 1))[FLAG_SORTDUP]) {
    s.busy = false;
    return fail(SMASH_ERR_DATA, "flags equal: %u lines of the batch share absolute position, name and first/second/reversed bits "
                "(MemSam::operator<, memsam.h:143-150, throws here: read names must be unique per mate)", ((const uint32_t *)(s.h_small.p + 1))[FLAG_SORTDUP]);
  }
  if (s.n_reads) {
    cudaEvent_t prev = s.ev0;
    for (int e = 0; e < s.n_evs; ++e) { float ms = 0; if (s.ev_stage[e] >= 0 && cudaEventElapsedTime(&ms, prev, s.evs[e]) == cudaSuccess) c->stage_ms[s.ev_stage[e]] += ms; prev = s.evs[e]; }
    s.n_evs = 0;
    if (g_dbg && s.tl[0] && g_tl_base && (s.want & SMASH_WANT_SAM)) {
      float a = 0, b0 = 0, d0 = 0, d1 = 0, e1 = 0;
      cudaEventElapsedTime(&a, g_tl_base, s.tl[0]); cudaEventElapsedTime(&b0, g_tl_base, s.ev0);
      cudaEventElapsedTime(&d0, g_tl_base, s.tl[1]); cudaEventElapsedTime(&d1, g_tl_base, s.tl[2]); cudaEventElapsedTime(&e1, g_tl_base, s.ev1);
      fprintf(stderr, "[smash-tl] slot %d  h2d %.2f..%.2f  compute ..%.2f  d2h %.2f..%.2f  end %.2f\n", (int)(&s - c->slot), a, b0, d0, d0, d1, e1);
    }
  }
  if (res) {
    memset(res, 0, sizeof *res);
    res->n_reads = s.n_reads; res->sam_bytes = s.sam_bytes; res->n_records = s.n_records;
    if (s.n_reads) { float ms = 0; cudaEventElapsedTime(&ms, s.ev0, s.ev1); res->gpu_ms = ms; }
    if ((s.want & SMASH_WANT_SAM) && s.h_sam.p) res->sam = s.h_sam.p;
    if ((s.want & SMASH_WANT_MATCHES) && s.n_reads) {
      res->match_off = s.h_csr_off.p; res->matches = s.h_matches.p;
      res->n_matches = (uint64_t)s.h_csr_off.p[s.n_reads];
    }
  }
  s.busy = false;
  return 0;
}

// ------------------------------------------------------------------ device-side input stage

static const char *ing_err_text(uint32_t code) {
  switch (code) {
    case ING_FEW_FIELDS: return "SAM line with fewer than 11 fields (the reference would reuse the previous line's fields, query.cpp:640-642)";
    case ING_BAD_FLAG: return "SAM flag field is not an unsigned integer";
    case ING_LEN_MISMATCH: return "SEQ and QUAL lengths differ";
    case ING_FQ_AT: return "Fastq @ parse error";                 // fastqs_to_sam.cpp:74
    case ING_FQ_PLUS: return "Fastq + parse error";               // fastqs_to_sam.cpp:71
    case ING_FQ_NAME: return "Problem reading read name";         // fastqs_to_sam.cpp:57
    case ING_FQ_TRUNC: return "FASTQ record cut short by the end of the input";
    case ING_FQ_COLUMNS: return "FASTQ bases/errors line is not one token (the SAM columns would shift)";
    case ING_TOO_LONG: return "field longer than 2^31 bytes";
    default: return "unknown input error";
  }
}

// Raw text (host) -> packed batch in the slot's device buffers; the parse runs on the GPU.  Synchronises the
// slot's stream (sizes have to reach the host before buffers are sized), so the text buffers may be reused
// when it returns.  info->consumed: bytes of each text that went into this batch.
static int ingest_text(smash_ctx *c, Slot &s, const smash_text *t, smash_text_info *info) {
  const bool fastq = t->kind == SMASH_TEXT_FASTQ_PAIR;
  const int final = (t->flags & SMASH_TEXT_FINAL) ? 1 : 0;
  const int n_text = fastq ? 2 : 1;
  const int phase = (fastq && (t->flags & SMASH_TEXT_MATE2_FIRST)) ? 1 : 0;
  cudaStream_t st = s.st;
  int rc;
  uint64_t nb[2] = {0, 0};
  for (int f = 0; f < n_text; ++f) {
    nb[f] = t->n_bytes[f];
    if (nb[f] && !t->text[f]) return fail(SMASH_ERR_ARG, "null text");
    if (!final) {                                            // a non-final chunk ends at its last complete line
      while (nb[f] && t->text[f][nb[f] - 1] != '\n') --nb[f];
    }
  }
  if (info) { info->n_reads = 0; info->consumed[0] = info->consumed[1] = 0; info->mate2_first_next = phase; }
  s.first_pair = t->first_pair_ordinal;
  if ((rc = s.h_ing.ensure(16)) || (rc = s.ing_scal.ensure(4))) return rc;
  uint64_t n_lines[2] = {0, 0};
  // 1. upload + count lines
  const uint64_t tiles[2] = {ing_tiles((nb[0] + 15) / 16), ing_tiles((nb[1] + 15) / 16)};
  if ((rc = s.ing_blk64.ensure(tiles[0] + tiles[1] + 4))) return rc;
  uint64_t *blk_lines[2] = {s.ing_blk64.p, s.ing_blk64.p + tiles[0] + 2};
  for (int f = 0; f < n_text; ++f) {
    if (!nb[f]) continue;
    if ((rc = s.ing_raw[f].ensure(nb[f] + 64))) return rc;
    CU(cudaMemcpyAsync(s.ing_raw[f].p, t->text[f], nb[f], cudaMemcpyHostToDevice, st));
  }
  if (!s.ev_ing0) { CU(cudaEventCreate(&s.ev_ing0)); CU(cudaEventCreate(&s.ev_ing1)); }
  CU(cudaEventRecord(s.ev_ing0, st));
  s.ing_timed = false;
  for (int f = 0; f < n_text; ++f) {
    if (!nb[f]) continue;
    c->launches += launch_ing_count_lines(s.ing_raw[f].p, nb[f], blk_lines[f], st);
    CU(cudaMemcpyAsync(s.h_ing.p + 8 + f, blk_lines[f] + tiles[f], 8, cudaMemcpyDeviceToHost, st));
  }
  if (nb[0] || nb[1]) CU(cudaStreamSynchronize(st));
  for (int f = 0; f < n_text; ++f) n_lines[f] = nb[f] ? s.h_ing.p[8 + f] : 0;
  // 2. line starts (+ FASTQ: header lines)
  CU(cudaMemsetAsync(s.ing_scal.p, 0xff, 8, st));
  CU(cudaMemsetAsync(s.ing_scal.p + 1, 0, 16, st));
  for (int f = 0; f < n_text; ++f) {
    if (!n_lines[f]) continue;
    if ((rc = s.ing_ls[f].ensure(n_lines[f] + 2))) return rc;
    c->launches += launch_ing_line_starts(s.ing_raw[f].p, nb[f], blk_lines[f], s.ing_ls[f].p, st);
  }
  uint64_t m = 0, n_rec[2] = {n_lines[0], 0}, n_take[2] = {0, 0};
  if (fastq) {
    const uint64_t max_lines = n_lines[0] > n_lines[1] ? n_lines[0] : n_lines[1];
    if (max_lines && ((rc = s.ing_hdr_flag.ensure(max_lines + 1)) || (rc = s.ing_blk32.ensure(ing_tiles(max_lines) + 2)))) return rc;
    for (int f = 0; f < 2; ++f) {
      if (!n_lines[f]) continue;
      if ((rc = s.ing_hdr[f].ensure(n_lines[f] + ing_tiles(n_lines[f]) + 4))) return rc;
      uint64_t *hdr = s.ing_hdr[f].p, *blk = s.ing_hdr[f].p + n_lines[f] + 1;    // record -> line table, then its scan tiles
      c->launches += launch_ing_fastq_headers(s.ing_raw[f].p, s.ing_ls[f].p, n_lines[f], s.ing_blk32.p, blk, s.ing_hdr_flag.p, hdr,
                                              (uint64_t *)(s.ing_scal.p + 1 + f), st);
    }
    CU(cudaMemcpyAsync(s.h_ing.p + 10, s.ing_scal.p + 1, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    n_rec[0] = s.h_ing.p[10]; n_rec[1] = s.h_ing.p[11];
    ing_fastq_take(n_rec[phase], n_rec[phase ^ 1], final, &n_take[phase], &n_take[phase ^ 1]);
    m = n_take[0] + n_take[1];
  } else {
    m = n_lines[0];
  }
  // 3. one LineRec per line / record, ordered compaction
  uint64_t n_reads = 0, name_bytes = 0, seq_bytes = 0, opt_bytes = 0, m_used = 0;
  if (m) {
    if ((rc = s.ing_recs.ensure(m + 1)) || (rc = s.ing_pre.ensure(m + 2)) || (rc = s.ing_blk4.ensure(ing_tiles(m) + 2))) return rc;
    if (fastq) {
      const int replace_n = (t->flags & SMASH_TEXT_REPLACE_N) ? 1 : 0;
      for (int f = 0; f < 2; ++f)
        c->launches += launch_ing_parse_fastq(s.ing_raw[f].p, s.ing_ls[f].p, n_lines[f], s.ing_hdr[f].p, n_take[f], f, phase, replace_n,
                                              s.ing_recs.p, s.ing_scal.p, st);
    } else {
      c->launches += launch_ing_parse_sam(s.ing_raw[0].p, s.ing_ls[0].p, n_lines[0], s.ing_recs.p, s.ing_scal.p, st);
    }
    c->launches += launch_ing_scan_recs(s.ing_recs.p, m, s.ing_blk4.p, s.ing_pre.p, st);
    IngPublish pb{};
    pb.pre = s.ing_pre.p; pb.m = m; pb.final = final; pb.fastq = fastq ? 1 : 0; pb.phase = phase;
    for (int f = 0; f < 2; ++f) { pb.ls[f] = s.ing_ls[f].p; pb.hdr[f] = s.ing_hdr[f].p; pb.n_rec[f] = n_rec[f]; pb.n_bytes[f] = nb[f]; }
    pb.err = s.ing_scal.p; pb.host = s.h_ing.p;
    c->launches += launch_ing_publish(pb, st);
    CU(cudaStreamSynchronize(st));
    const uint64_t err = s.h_ing.p[7];
    if (err != ~0ull) {
      const uint32_t code = (uint32_t)(err & 0xff);
      const uint64_t idx = err >> 8;
      if (fastq) return fail(SMASH_ERR_DATA, "%s (record %llu of mate file %d in this chunk)", ing_err_text(code), (unsigned long long)(idx / 2 + 1), (int)((idx & 1) ^ phase) + 1);
      return fail(SMASH_ERR_DATA, "%s (input line %llu of this chunk)", ing_err_text(code), (unsigned long long)(idx + 1));
    }
    n_reads = s.h_ing.p[0]; name_bytes = s.h_ing.p[1]; seq_bytes = s.h_ing.p[2]; opt_bytes = s.h_ing.p[3]; m_used = s.h_ing.p[4];
    if (info) { info->consumed[0] = s.h_ing.p[5]; info->consumed[1] = s.h_ing.p[6]; info->mate2_first_next = (int)s.h_ing.p[12]; }
  } else if (info) {
    // nothing to parse: a final call consumes what is left, a non-final one waits for more input
    info->consumed[0] = final ? t->n_bytes[0] : 0; info->consumed[1] = (final && fastq) ? t->n_bytes[1] : 0;
  }
  if (final && info) { info->consumed[0] = t->n_bytes[0]; info->consumed[1] = fastq ? t->n_bytes[1] : 0; }
  // 4. batch buffers, then the copy
  if ((rc = slot_reserve(c, s, n_reads, name_bytes, seq_bytes, opt_bytes, 1))) return rc;
  if (n_reads) {
    IngCopy cp{};
    cp.text[0] = s.ing_raw[0].p; cp.text[1] = s.ing_raw[1].p; cp.recs = s.ing_recs.p; cp.pre = s.ing_pre.p; cp.m = m_used;
    cp.names = s.names.p; cp.name_off = s.name_off.p; cp.seq = s.seq.p; cp.qual = s.qual.p; cp.seq_off = s.seq_off.p;
    cp.opt = opt_bytes ? s.opt.p : nullptr; cp.opt_off = opt_bytes ? s.opt_off.p : nullptr; cp.read_flag = s.read_flag.p;
    c->launches += launch_ing_copy(cp, st);
  }
  CU(cudaEventRecord(s.ev_ing1, st));
  s.ing_timed = true;
  if (info) info->n_reads = n_reads;
  return 0;
}

extern "C" int smash_submit_text(smash_ctx *c, int slot, const smash_text *t, int want, smash_text_info *info) {
  if (!c || !t || slot < 0 || slot >= SMASH_N_SLOTS) return fail(SMASH_ERR_ARG, "bad argument");
  if (t->kind != SMASH_TEXT_SAM && t->kind != SMASH_TEXT_FASTQ_PAIR) return fail(SMASH_ERR_ARG, "unknown text kind %d", t->kind);
  Slot &s = c->slot[slot];
  if (s.busy) return fail(SMASH_ERR_STATE, "slot %d still has a batch in flight", slot);
  CU(cudaSetDevice(c->device));
  s.compact = false;                                         // the packed batch exists only on the device: full SAM text comes back
  int rc = ingest_text(c, s, t, info);
  if (rc) return rc;
  s.job_prepare = false; s.job_want = want; s.job_chunks = 1;
  return job_start(c, s);
}
extern "C" int smash_text_upload(smash_ctx *c, const smash_text *t, smash_text_info *info) {
  if (!c || !t) return fail(SMASH_ERR_ARG, "null argument");
  if (t->kind != SMASH_TEXT_SAM && t->kind != SMASH_TEXT_FASTQ_PAIR) return fail(SMASH_ERR_ARG, "unknown text kind %d", t->kind);
  CU(cudaSetDevice(c->device));
  Slot &s = c->slot[0];
  if (s.busy) return fail(SMASH_ERR_STATE, "slot 0 still has a batch in flight");
  int rc = ingest_text(c, s, t, info);
  if (rc) return rc;
  CU(cudaStreamSynchronize(s.st));
  CU(cudaGetLastError());
  ingest_collect(c, s);
  return 0;
}
extern "C" double smash_ctx_ingest_ms(smash_ctx *c, int reset) {
  if (!c) return 0.0;
  const double v = c->ingest_ms;
  if (reset) c->ingest_ms = 0;
  return v;
}
extern "C" int smash_batch_sizes(smash_ctx *c, int slot, uint64_t *n_reads, uint64_t *name_bytes, uint64_t *seq_bytes, uint64_t *opt_bytes) {
  if (!c || slot < 0 || slot >= SMASH_N_SLOTS) return fail(SMASH_ERR_ARG, "bad argument");
  const Slot &s = c->slot[slot];
  if (n_reads) *n_reads = s.n_reads;
  if (name_bytes) *name_bytes = s.name_bytes;
  if (seq_bytes) *seq_bytes = s.seq_bytes;
  if (opt_bytes) *opt_bytes = s.opt_bytes;
  return 0;
}
extern "C" int smash_fetch_batch(smash_ctx *c, int slot, uint8_t *names, int64_t *name_off, uint8_t *seq, uint8_t *qual, int64_t *seq_off,
                                 uint8_t *opt, int64_t *opt_off, uint16_t *read_flag) {
  if (!c || slot < 0 || slot >= SMASH_N_SLOTS) return fail(SMASH_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  Slot &s = c->slot[slot];
  CU(cudaStreamSynchronize(s.st));
  const uint64_t n = s.n_reads;
  if (!n) return 0;
  if (names) CU(cudaMemcpy(names, s.names.p, s.name_bytes, cudaMemcpyDeviceToHost));
  if (name_off) CU(cudaMemcpy(name_off, s.name_off.p, 8 * (n + 1), cudaMemcpyDeviceToHost));
  if (seq) CU(cudaMemcpy(seq, s.seq.p, s.seq_bytes, cudaMemcpyDeviceToHost));
  if (qual) CU(cudaMemcpy(qual, s.qual.p, s.seq_bytes, cudaMemcpyDeviceToHost));
  if (seq_off) CU(cudaMemcpy(seq_off, s.seq_off.p, 8 * (n + 1), cudaMemcpyDeviceToHost));
  if (s.opt_bytes) {
    if (opt) CU(cudaMemcpy(opt, s.opt.p, s.opt_bytes, cudaMemcpyDeviceToHost));
    if (opt_off) CU(cudaMemcpy(opt_off, s.opt_off.p, 8 * (n + 1), cudaMemcpyDeviceToHost));
  } else if (opt_off) {
    memset(opt_off, 0, 8 * (n + 1));
  }
  if (read_flag) CU(cudaMemcpy(read_flag, s.read_flag.p, 2 * n, cudaMemcpyDeviceToHost));
  return 0;
}

// One batch on its slot's worker thread: H2D enqueue, the per-range kernel launches (each reads its sizes back before
// the emit kernels), the tail appends in submission order, and -- compact transport -- the expansion into SAM lines.
static int run_job(smash_ctx *c, Slot &s) {
  int rc = 0;
  const double t0 = now_ms();
  s.expand_ns = 0;
  if (s.job_prepare) rc = slot_prepare(c, s, &s.hb, true, s.job_chunks);
  const double t1 = now_ms();
  if (!rc) rc = slot_run(c, s, s.job_want, true);
  const double t2 = now_ms();
  tail_turn_release(c, s);                                   // the next batch may append now
  if (s.compact) {
    if (!rc) rc = expand_poll(c, s, true);
    const double t3 = now_ms();
    expand_drain(c, s);                                      // also on errors: tasks reference the slot's buffers
    if (g_dbg) fprintf(stderr, "[smash-dbg] job slot %d @%.2f: prepare %.2f  run %.2f  d2h-wait %.2f  drain %.2f  (task time %.2f ms over %zu threads)\n",
                       (int)(&s - c->slot), t0, t1 - t0, t2 - t1, t3 - t2, now_ms() - t3, (double)s.expand_ns.load() / 1e6, c->pool.th.size());
  } else if (g_dbg) fprintf(stderr, "[smash-dbg] job slot %d @%.2f: prepare %.2f  run %.2f\n", (int)(&s - c->slot), t0, t1 - t0, t2 - t1);
  return rc;
}
static void slot_worker_main(smash_ctx *c, Slot *sp) {
  Slot &s = *sp;
  cudaSetDevice(c->device);
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(s.mu);
      s.cv.wait(lk, [&] { return s.quit || s.job_state == 1; });
      if (s.quit) return;
      s.job_state = 2;
    }
    const int rc = run_job(c, s);
    if (rc) { strncpy(s.job_err, g_err, sizeof s.job_err - 1); s.job_err[sizeof s.job_err - 1] = 0; }
    { std::lock_guard<std::mutex> lk(s.mu); s.job_rc = rc; s.job_state = 3; }
    s.cv.notify_all();
  }
}
static int job_start(smash_ctx *c, Slot &s) {
  slot_begin_job(c, s);
  s.busy = true;
#if defined(SMASH_CUDA_SHIM)
  s.job_rc = run_job(c, s);                                  // host emulation (tests/emul): no threads
  if (s.job_rc) { strncpy(s.job_err, g_err, sizeof s.job_err - 1); s.job_err[sizeof s.job_err - 1] = 0; }
  s.job_state = 3;
#else
  if (s.compact) pool_start(c);
  if (!s.worker.joinable()) s.worker = std::thread(slot_worker_main, c, &s);
  { std::lock_guard<std::mutex> lk(s.mu); s.job_state = 1; }
  s.cv.notify_all();
#endif
  return 0;
}

extern "C" int smash_submit(smash_ctx *c, int slot, const smash_batch *b, int want) {
  if (!c || !b || slot < 0 || slot >= SMASH_N_SLOTS) return fail(SMASH_ERR_ARG, "bad argument");
  Slot &s = c->slot[slot];
  if (s.busy) return fail(SMASH_ERR_STATE, "slot %d still has a batch in flight", slot);
  if (b->n_reads && (!b->names || !b->name_off || !b->seq || !b->qual || !b->seq_off || !b->read_flag)) return fail(SMASH_ERR_ARG, "batch with null arrays");
  CU(cudaSetDevice(c->device));
  if (g_dbg) {
    if (!g_tl_base) { cudaEventCreate(&g_tl_base); cudaEventRecord(g_tl_base, s.st); }
    for (int e = 0; e < 3; ++e) if (!s.tl[e]) cudaEventCreate(&s.tl[e]);
    cudaEventRecord(s.tl[0], s.st);
  }
  // (a sorted batch is ONE chunk of the reference's OutputSorter: it goes through whole)
  s.job_chunks = (c->prm.mode != SMASH_MODE_MEM && !(want & (SMASH_WANT_MATCHES | SMASH_WANT_SORTED)) && (want & SMASH_WANT_SAM) && !g_no_chunks) ? c->max_chunks : 1;
  s.hb = *b; s.job_prepare = true; s.job_want = want;
  // the caller's batch stays valid until smash_wait, so the lines can be rebuilt from it on the host: compact transport
  s.compact = (want & SMASH_WANT_SAM) && !g_full_sam && c->transport != 1;
  s.ctx = c;
  return job_start(c, s);
}
extern "C" int smash_wait(smash_ctx *c, int slot, smash_result *res) {
  if (!c || slot < 0 || slot >= SMASH_N_SLOTS) return fail(SMASH_ERR_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  const double t0 = now_ms();
  Slot &s = c->slot[slot];
  if (s.job_state != 0) {
    {
      std::unique_lock<std::mutex> lk(s.mu);
      s.cv.wait(lk, [&] { return s.job_state == 3; });
      s.job_state = 0;
    }
    if (s.job_rc) {
      cudaStreamSynchronize(s.st); cudaGetLastError();
      c->launches += s.launches; s.launches = 0; s.io_h2d = 0; s.io_d2h = 0;
      s.busy = false;
      return fail(s.job_rc, "%s", s.job_err);
    }
  }
  const int rc = slot_finish(c, s, res);
  DBG_T("wait", t0);
  return rc;
}
extern "C" int smash_map_batch(smash_ctx *c, const smash_batch *b, int want, smash_result *res) {
  int rc = smash_submit(c, 0, b, want);
  if (rc) return rc;
  return smash_wait(c, 0, res);
}

extern "C" int smash_batch_upload(smash_ctx *c, const smash_batch *b) {
  if (!c || !b) return fail(SMASH_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  Slot &s = c->slot[0];
  if (s.busy) return fail(SMASH_ERR_STATE, "slot 0 still has a batch in flight");
  s.compact = false;
  int rc = slot_prepare(c, s, b, true);
  if (rc) return rc;
  CU(cudaStreamSynchronize(s.st));
  c->io_h2d += s.io_h2d; s.io_h2d = 0;
  return 0;
}
extern "C" int smash_map_resident(smash_ctx *c, int want, smash_result *res) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  Slot &s = c->slot[0];
  if (s.busy) return fail(SMASH_ERR_STATE, "slot 0 still has a batch in flight");
  s.compact = false;
  slot_begin_job(c, s);
  int rc = slot_run(c, s, want & ~SMASH_WANT_MATCHES, false);
  tail_turn_release(c, s);
  if (rc) return rc;
  rc = slot_finish(c, s, res);
  if (res) res->sam = nullptr;
  return rc;
}
extern "C" int smash_fetch_sam(smash_ctx *c, const char **sam, uint64_t *n_bytes) {
  if (!c || !sam || !n_bytes) return fail(SMASH_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  Slot &s = c->slot[0];
  int rc = s.h_sam.ensure(s.sam_bytes + 64);
  if (rc) return rc;
  CU(cudaMemcpy(s.h_sam.p, s.sam.p, s.sam_bytes, cudaMemcpyDeviceToHost));
  *sam = s.h_sam.p; *n_bytes = s.sam_bytes;
  return 0;
}

// ------------------------------------------------------------------ tail + misc

// the tail calls below read or reset what the slot workers append to
static int require_idle(smash_ctx *c) {
  for (int i = 0; i < SMASH_N_SLOTS; ++i)
    if (c->slot[i].busy) return fail(SMASH_ERR_STATE, "slot %d still has a batch in flight: collect it with smash_wait first", i);
  return 0;
}
// ---- accessors for comm.cu (ctx_internal.h)
int ctx_fail(int code, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return code;
}
void **ctx_comm_slot(smash_ctx *c) { return &c->comm; }
int ctx_device(const smash_ctx *c) { return c->device; }
smash::TailState *ctx_tail(smash_ctx *c) { return &c->tail; }
cudaStream_t ctx_stream(smash_ctx *c) { return c->slot[0].st; }
uint64_t *ctx_launches(smash_ctx *c) { return &c->launches; }
uint64_t ctx_n_bins(smash_ctx *c) { return c->tail.n_bins; }
int ctx_require_idle(smash_ctx *c) { return require_idle(c); }

extern "C" int smash_tail_configure(smash_ctx *c, const int64_t *bin_starts, uint64_t n_bins,
                                    const char *const *chrom_names, const int64_t *chrom_offsets,
                                    uint64_t n_chroms, int64_t hit_window, int32_t min_excess) {
  if (!c || !bin_starts || !n_bins) return fail(SMASH_ERR_ARG, "bad argument");
  if (!c->dix.mapbody) return fail(SMASH_ERR_STATE, "the tail needs map.bin (smash_ctx_load_mappability / smash_ctx_build_mappability)");
  if (!c->hix->rcref) return fail(SMASH_ERR_ARG, "the tail requires an -rcref index (forward sequences at even seq_index)");
  CU(cudaSetDevice(c->device));
  // per forward chromosome: passes /^chr(\d+|[XY])$/ (smash_mapping.sh:29) and is listed in
  // chrom_sizes.txt without '_' / chrM (varbin.py:38-49) -> its absolute offset, else -1
  std::vector<int64_t> off;
  const int step = c->hix->rcref ? 2 : 1;
  for (size_t i = 0; i < c->hix->descr.size(); i += step) {
    const std::string &nm = c->hix->descr[i];
    int64_t o = -1;
    bool re = nm.size() > 3 && nm.compare(0, 3, "chr") == 0;
    if (re) {
      const std::string t = nm.substr(3);
      bool digits = !t.empty();
      for (char ch : t) digits = digits && ch >= '0' && ch <= '9';
      re = digits || t == "X" || t == "Y";
    }
    if (re && nm.find('_') == std::string::npos && nm != "chrM")
      for (uint64_t k = 0; k < n_chroms; ++k)
        if (nm == chrom_names[k]) { o = chrom_offsets[k]; break; }
    off.push_back(o);
  }
  int rc = tail_configure(&c->tail, bin_starts, n_bins, off.data(), off.size(), hit_window, min_excess);
  if (rc) return fail(rc, "tail: %s", tail_error());
  return 0;
}
extern "C" int smash_tail_finish(smash_ctx *c, int64_t *counts, void *counts_device, smash_tail_stats *st) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  if (int idle_rc = require_idle(c)) return idle_rc;
  CU(cudaSetDevice(c->device));
  for (int s = 0; s < SMASH_N_SLOTS; ++s) CU(cudaStreamSynchronize(c->slot[s].st));
  const double tf = now_ms();
  int rc = tail_finish(&c->tail, counts, (int64_t *)counts_device, st, c->slot[0].st, &c->launches);
  DBG_T("tail_finish", tf);
  if (rc) return fail(rc, "tail: %s", tail_error());
  return 0;
}
extern "C" int smash_tail_export_keys(smash_ctx *c, uint64_t ordinal_base, const void **dev_keys, uint64_t *n_keys) {
  if (!c || !dev_keys || !n_keys) return fail(SMASH_ERR_ARG, "null argument");
  if (int idle_rc = require_idle(c)) return idle_rc;
  CU(cudaSetDevice(c->device));
  for (int s = 0; s < SMASH_N_SLOTS; ++s) CU(cudaStreamSynchronize(c->slot[s].st));
  const uint64_t *k = nullptr;
  int rc = tail_export_keys(&c->tail, ordinal_base, &k, n_keys, c->slot[0].st, &c->launches);
  if (rc) return fail(rc, "tail: %s", tail_error());
  *dev_keys = k;
  return 0;
}
extern "C" int smash_tail_phase_a(smash_ctx *c, uint64_t ordinal_base, const void *foreign_keys_dev, uint64_t n_foreign,
                                  smash_tail_edge *edge) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  if (int idle_rc = require_idle(c)) return idle_rc;
  CU(cudaSetDevice(c->device));
  for (int s = 0; s < SMASH_N_SLOTS; ++s) CU(cudaStreamSynchronize(c->slot[s].st));
  int rc = tail_phase_a(&c->tail, ordinal_base, (const uint64_t *)foreign_keys_dev, n_foreign, edge, c->slot[0].st, &c->launches, nullptr, true);
  if (rc) return fail(rc, "tail: %s", tail_error());
  return 0;
}
extern "C" int smash_tail_phase_a_verdict(smash_ctx *c, uint64_t ordinal_base, const void *min_ordinal_dev, uint64_t n_keys,
                                          smash_tail_edge *edge) {
  if (!c || (!min_ordinal_dev && n_keys)) return fail(SMASH_ERR_ARG, "null argument");
  if (int idle_rc = require_idle(c)) return idle_rc;
  CU(cudaSetDevice(c->device));
  for (int s = 0; s < SMASH_N_SLOTS; ++s) CU(cudaStreamSynchronize(c->slot[s].st));
  static const uint64_t dummy = 0;
  int rc = tail_phase_a(&c->tail, ordinal_base, nullptr, 0, edge, c->slot[0].st, &c->launches,
                        n_keys ? (const uint64_t *)min_ordinal_dev : &dummy, true);
  if (rc) return fail(rc, "tail: %s", tail_error());
  return 0;
}
extern "C" int smash_tail_phase_b(smash_ctx *c, int has_prev, int64_t prev_last_pos, int64_t *counts, void *counts_device,
                                  smash_tail_stats *st) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  if (int idle_rc = require_idle(c)) return idle_rc;
  CU(cudaSetDevice(c->device));
  int rc = tail_phase_b(&c->tail, has_prev, prev_last_pos, counts, (int64_t *)counts_device, st, c->slot[0].st, &c->launches);
  if (rc) return fail(rc, "tail: %s", tail_error());
  return 0;
}
extern "C" int smash_tail_positions(smash_ctx *c, const int32_t **chrom, const int64_t **pos, uint64_t *n) {
  if (!c || !chrom || !pos || !n) return fail(SMASH_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  int rc = tail_positions(&c->tail, chrom, pos, n);
  if (rc) return fail(rc, "tail: %s", tail_error());
  return 0;
}
extern "C" int smash_tail_reserve(smash_ctx *c, uint64_t max_pairs, uint64_t max_hits) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  int rc = tail_reserve(&c->tail, max_pairs, max_hits, c->slot[0].st);
  if (rc) return fail(rc, "tail: %s", tail_error());
  return 0;
}
extern "C" int smash_tail_reset(smash_ctx *c) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  if (int idle_rc = require_idle(c)) return idle_rc;
  CU(cudaSetDevice(c->device));
  tail_reset(&c->tail);
  return 0;
}

// Release the inverse suffix array (49.5 GB at hg19 scale) once map.bin is built / the index is saved;
// MEM mode needs it and keeps it.
extern "C" int smash_ctx_drop_isa(smash_ctx *c) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  if (c->prm.mode == SMASH_MODE_MEM) return fail(SMASH_ERR_STATE, "MEM mode needs the inverse suffix array");
  CU(cudaSetDevice(c->device));
  if (c->isa) { cudaFree(c->isa); c->index_bytes -= c->dix.N * (uint64_t)c->dix.w; c->isa = nullptr; c->dix.isa = nullptr; }
  return 0;
}
extern "C" int smash_memcpy(void *dst, const void *src, size_t bytes) {
  if (!bytes) return 0;
  CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));       // UVA: any of host/device on either side
  return 0;
}
extern "C" int smash_ctx_set_tag_mappability(smash_ctx *c, int on) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  if (on && !c->dix.mapbody) return fail(SMASH_ERR_STATE, "tagging needs map.bin (smash_ctx_load_mappability / smash_ctx_build_mappability)");
  for (int s = 0; s < SMASH_N_SLOTS; ++s) if (c->slot[s].busy) return fail(SMASH_ERR_STATE, "slot %d still has a batch in flight", s);
  c->prm.tag_mappability = on ? 1 : 0; c->sp.tag_mappability = c->prm.tag_mappability;
  return 0;
}
extern "C" int smash_ctx_set_chunking(smash_ctx *c, int max_chunks, uint64_t min_reads) {
  if (!c || max_chunks < 1 || max_chunks > MAX_CHUNKS) return fail(SMASH_ERR_ARG, "max_chunks must be 1..%d", MAX_CHUNKS);
  c->max_chunks = max_chunks; c->chunk_min_reads = min_reads < 2 ? 2 : min_reads;
  return 0;
}
extern "C" int smash_ctx_set_transport(smash_ctx *c, int full_sam_text, int host_threads) {
  if (!c) return fail(SMASH_ERR_ARG, "null argument");
  for (int i = 0; i < SMASH_N_SLOTS; ++i) if (c->slot[i].busy) return fail(SMASH_ERR_STATE, "slot %d has a batch in flight", i);
  if (full_sam_text < 0 || full_sam_text > 3) return fail(SMASH_ERR_ARG, "transport %d: 0 = per range whichever is faster, 1 = full SAM text, 2 = compact only", full_sam_text);
  c->transport = full_sam_text;
  if (host_threads != c->host_threads && !c->pool.th.empty()) {       // restart the expansion pool with the new size
    { std::lock_guard<std::mutex> lk(c->pool.mu); c->pool.quit = true; }
    c->pool.cv.notify_all();
    for (auto &t : c->pool.th) t.join();
    c->pool.th.clear(); c->pool.quit = false;
  }
  c->host_threads = host_threads;
  return 0;
}
extern "C" void smash_ctx_io_bytes(smash_ctx *c, uint64_t *h2d, uint64_t *d2h, int reset) {
  if (!c) return;
  if (h2d) *h2d = c->io_h2d;
  if (d2h) *d2h = c->io_d2h;
  if (reset) { c->io_h2d = 0; c->io_d2h = 0; }
}
extern "C" int smash_host_expand(const smash_batch *b, uint64_t read_base, const void *meta, uint64_t n_records, const char *cmp, char *sam) {
  if (!b || !meta || !cmp || !sam) return fail(SMASH_ERR_ARG, "null argument");
  ExpandArgs a{};
  a.names = b->names; a.name_off = b->name_off; a.seq = b->seq; a.qual = b->qual; a.seq_off = b->seq_off;
  a.opt = b->opt; a.opt_off = b->opt ? b->opt_off : nullptr;
  a.read_base = read_base; a.meta = (const CmpMeta *)meta; a.cmp = cmp; a.sam = sam;
  expand_records(a, 0, n_records);
  return 0;
}
extern "C" uint64_t smash_ctx_launch_count(const smash_ctx *c) { return c ? c->launches : 0; }
extern "C" void smash_ctx_stage_ms(smash_ctx *c, double *out, int reset) {
  if (!c) return;
  if (out) for (int i = 0; i < 8; ++i) out[i] = c->stage_ms[i];
  if (reset) for (int i = 0; i < 8; ++i) c->stage_ms[i] = 0;
}
extern "C" uint64_t smash_ctx_index_bytes(const smash_ctx *c) { return c ? c->index_bytes : 0; }
extern "C" void *smash_ctx_stream(const smash_ctx *c) { return c ? (void *)c->slot[0].st : nullptr; }
