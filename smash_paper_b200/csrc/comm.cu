// comm.cu -- the multi-GPU half of the tail behind the C ABI: smash_comm_* and smash_bins_finish.
//
// Reads shard across GPUs as contiguous ranges of pairs, the index is replicated, nothing is exchanged on the search /
// SAM path.  Two order-dependent rules of the tail cross shard boundaries and are resolved here, on the
// device, with NCCL moving the data:
//   1. smashMEM.py's first-wins duplicate rule (smashMEM.py:217-228) is global.  The key space is hash-partitioned:
//      every rank buckets its dupe-set keys {fp1, fp2, ordinal} by fp1 mod world (k_owner_count / k_owner_scatter), one
//      all-to-all (grouped ncclSend/ncclRecv) moves each bucket to its owner, the owner radix-sorts what it got by fp1
//      and gives every row the smallest ordinal of its (fp1, fp2) group (k_group_min), a second all-to-all returns the
//      verdicts; per-rank work and traffic do not grow with the number of ranks.
//   2. varbin.py's "same position as the previous kept line" rule (varbin.py:56-58): an all-gather of the shard edges.
// Then ONE ncclAllReduce(sum) of the n_bins int64 counts (+ the six stats counters).
// NCCL is bound at run time (dlopen): the library loads and maps on a single GPU without it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/smash_b200.h"
#include "ctx_internal.h"
#include "tail.cuh"

using namespace smash;

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.lib) return 0;
  void *h = nullptr;
  if (const char *e = getenv("SMASH_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // the copy the host program already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return ctx_fail(SMASH_ERR_STATE, "NCCL not found (libnccl.so.2; set SMASH_NCCL_LIB): %s", dlerror());
#define BIND(name) do { *(void **)(&g_nccl.name) = dlsym(h, "nccl" #name); if (!g_nccl.name) return ctx_fail(SMASH_ERR_STATE, "NCCL symbol nccl" #name " missing"); } while (0)
  BIND(GetUniqueId); BIND(CommInitRank); BIND(CommInitAll); BIND(CommDestroy); BIND(AllGather); BIND(AllReduce);
  BIND(Send); BIND(Recv); BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString);
#undef BIND
  g_nccl.lib = h;
  return 0;
}
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return ctx_fail(SMASH_ERR_CUDA, "%s: %s", #call, g_nccl.GetErrorString(r_)); } while (0)
#define CC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(SMASH_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

template <class T> struct Buf {                                // growable device scratch
  T *p = nullptr; size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = n + n / 4 + 64;
    if (cudaMalloc((void **)&p, want * sizeof(T)) != cudaSuccess) return ctx_fail(SMASH_ERR_NOMEM, "cudaMalloc(%zu bytes)", want * sizeof(T));
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct CommState {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  Buf<uint64_t> send, recv, fp1, fp1_sorted, mins, back, verdict, small, counts, info;
  Buf<uint32_t> perm, idx, idx_sorted;
  Buf<uint8_t> sort_tmp;
  uint64_t *h_small = nullptr;                                 // pinned: counts matrix, edges, stats
};

// ---- kernels ---------------------------------------------------------------------------------------------------------

constexpr int MAX_WORLD = 64;

__global__ void k_owner_count(const uint64_t *__restrict__ keys, uint64_t n, int world, uint64_t *__restrict__ cnt) {
  __shared__ unsigned int loc[MAX_WORLD];
  for (int i = threadIdx.x; i < world; i += blockDim.x) loc[i] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(&loc[keys[3 * i] % (uint64_t)world], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < world; i += blockDim.x) if (loc[i]) atomicAdd((unsigned long long *)&cnt[i], (unsigned long long)loc[i]);
}
// rows go to their owner's bucket of the send buffer; perm[i] = where row i went (the verdicts come back in place)
__global__ void k_owner_scatter(const uint64_t *__restrict__ keys, uint64_t n, int world, const uint64_t *__restrict__ off,
                                uint64_t *__restrict__ cursor, uint64_t *__restrict__ send, uint32_t *__restrict__ perm) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t fp1 = keys[3 * i], fp2 = keys[3 * i + 1], ord = keys[3 * i + 2];
    const int o = (int)(fp1 % (uint64_t)world);
    const uint64_t pos = off[o] + (uint64_t)atomicAdd((unsigned long long *)&cursor[o], 1ull);
    send[3 * pos] = fp1; send[3 * pos + 1] = fp2; send[3 * pos + 2] = ord;
    perm[i] = (uint32_t)pos;
  }
}
__global__ void k_take_fp1(const uint64_t *__restrict__ rows, uint64_t n, uint64_t *__restrict__ fp1, uint32_t *__restrict__ idx) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) { fp1[i] = rows[3 * i]; idx[i] = (uint32_t)i; }
}
// rows sorted by fp1 (idx_sorted): the head of every run of equal fp1 gives each row of the run the smallest ordinal
// among the run's rows with the same fp2 (runs hold one row unless pairs really are duplicates)
__global__ void k_group_min(const uint64_t *__restrict__ rows, const uint64_t *__restrict__ fp1_sorted, const uint32_t *__restrict__ idx_sorted,
                            uint64_t n, uint64_t *__restrict__ mins) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    if (j && fp1_sorted[j - 1] == fp1_sorted[j]) continue;
    uint64_t e = j + 1;
    while (e < n && fp1_sorted[e] == fp1_sorted[j]) ++e;
    for (uint64_t t = j; t < e; ++t) {
      const uint32_t it = idx_sorted[t];
      const uint64_t fp2 = rows[3 * (uint64_t)it + 1];
      uint64_t m = rows[3 * (uint64_t)it + 2];
      for (uint64_t u = j; u < e; ++u) {
        const uint32_t iu = idx_sorted[u];
        if (rows[3 * (uint64_t)iu + 1] == fp2 && rows[3 * (uint64_t)iu + 2] < m) m = rows[3 * (uint64_t)iu + 2];
      }
      mins[it] = m;
    }
  }
}
__global__ void k_unpermute(const uint64_t *__restrict__ back, const uint32_t *__restrict__ perm, uint64_t n, uint64_t *__restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = back[perm[i]];
}
int grid_for(uint64_t n) { const uint64_t g = (n + 255) / 256; return (int)(g < 1 ? 1 : g > 148 * 8 ? 148 * 8 : g); }

// all-to-all of `width` uint64 per row: send_cnt[r] rows to rank r from send + width * send_off[r], recv likewise
int all_to_all(CommState *cs, const uint64_t *send, const uint64_t *send_cnt, const uint64_t *send_off, uint64_t *recv, const uint64_t *recv_cnt,
               const uint64_t *recv_off, int width, cudaStream_t st) {
  NC(g_nccl.GroupStart());
  for (int r = 0; r < cs->world; ++r) {
    if (send_cnt[r]) NC(g_nccl.Send(send + width * send_off[r], width * send_cnt[r], ncclUint64, r, cs->comm, st));
    if (recv_cnt[r]) NC(g_nccl.Recv(recv + width * recv_off[r], width * recv_cnt[r], ncclUint64, r, cs->comm, st));
  }
  NC(g_nccl.GroupEnd());
  return 0;
}

int comm_attach(smash_ctx *c, ncclComm_t comm, int rank, int world) {
  CommState *cs = new CommState();
  cs->comm = comm; cs->rank = rank; cs->world = world;
  if (cudaHostAlloc((void **)&cs->h_small, 8 * (size_t)(MAX_WORLD * MAX_WORLD + 64), cudaHostAllocDefault) != cudaSuccess) { delete cs; return ctx_fail(SMASH_ERR_NOMEM, "cudaHostAlloc"); }
  *ctx_comm_slot(c) = cs;
  return 0;
}

// ---- inputs that are not in `samtools sort -n` order -------------------------------------------------------------------
// smashMEM.py's first-wins rule and varbin.py's adjacent-duplicate rule follow the NAME order of the pairs, and the
// sharded protocol above relies on every shard being a contiguous run of it.  Every rank publishes {pairs, hits, name
// bytes, pairs out of order, first name, last name}; if any shard is out of order inside, or two neighbouring shards are
// out of order at their seam, the ranks send their pairs to rank 0 (grouped ncclSend/ncclRecv, rank order = input order)
// and rank 0 runs the single-GPU finish, whose name sort handles any order.  The decision is taken from all-gathered data,
// so every rank takes the same branch.
constexpr int INFO_WORDS = 4 + 2 * TAIL_EDGE_NAME / 8;
int gather_if_unsorted(CommState *cs, TailState *t, cudaStream_t st, uint64_t *launches, bool *gathered) {
  const int W = cs->world, R = cs->rank;
  int rc;
  std::vector<uint64_t> mine(INFO_WORDS), all((size_t)INFO_WORDS * W);
  if ((rc = tail_totals(t, mine.data(), st))) return ctx_fail(rc, "tail: %s", tail_error());
  if ((rc = tail_edge_names(t, (uint8_t *)(mine.data() + 4), (uint8_t *)(mine.data() + 4) + TAIL_EDGE_NAME, st))) return ctx_fail(rc, "tail: %s", tail_error());
  if ((rc = cs->info.ensure((size_t)INFO_WORDS * (W + 1)))) return rc;
  CC(cudaMemcpyAsync(cs->info.p, mine.data(), 8 * INFO_WORDS, cudaMemcpyHostToDevice, st));
  NC(g_nccl.AllGather(cs->info.p, cs->info.p + INFO_WORDS, INFO_WORDS, ncclUint64, cs->comm, st));
  CC(cudaMemcpyAsync(all.data(), cs->info.p + INFO_WORDS, 8 * (size_t)INFO_WORDS * W, cudaMemcpyDeviceToHost, st));
  CC(cudaStreamSynchronize(st));
  bool unsorted = false;
  const uint8_t *prev_last = nullptr;
  for (int r = 0; r < W; ++r) {
    const uint64_t *info = all.data() + (size_t)INFO_WORDS * r;
    if (!info[0]) continue;
    const uint8_t *first = (const uint8_t *)(info + 4), *last = first + TAIL_EDGE_NAME;
    if (info[3] || first[0] == 255 || last[0] == 255) unsorted = true;       // a name too long for the slot: take the safe path
    else if (prev_last && strnum_cmp(prev_last + 1, prev_last[0], first + 1, first[0]) > 0) unsorted = true;
    prev_last = last;
  }
  *gathered = unsorted;
  if (!unsorted) return 0;
  // rank 0 takes the others' pairs behind its own, shard by shard
  uint64_t add[3] = {0, 0, 0};
  for (int r = 1; r < W; ++r) for (int k = 0; k < 3; ++k) add[k] += all[(size_t)INFO_WORDS * r + k];
  if (R == 0 && (rc = tail_absorb_reserve(t, add[0], add[1], add[2], st))) return ctx_fail(rc, "tail: %s", tail_error());
  NC(g_nccl.GroupStart());
  if (R == 0) {
    uint64_t P = all[0], H = all[1], B = all[2];
    for (int r = 1; r < W; ++r) {
      const uint64_t *info = all.data() + (size_t)INFO_WORDS * r;
      const uint64_t p = info[0], h = info[1], b = info[2];
      if (p) {
        NC(g_nccl.Recv(t->pair_nhits.p + P, 4 * p, ncclUint8, r, cs->comm, st));
        NC(g_nccl.Recv(t->pair_fp.p + 2 * P, 16 * p, ncclUint8, r, cs->comm, st));
        NC(g_nccl.Recv(t->pair_hit_off.p + P, 8 * p, ncclUint8, r, cs->comm, st));
        NC(g_nccl.Recv(t->pair_name_off.p + P + 1, 8 * p, ncclUint8, r, cs->comm, st));
        if (h) NC(g_nccl.Recv(t->hits.p + H, 8 * h, ncclUint8, r, cs->comm, st));
        if (b) NC(g_nccl.Recv(t->name_blob.p + B, b, ncclUint8, r, cs->comm, st));
      }
      P += p; H += h; B += b;
    }
  } else if (mine[0]) {
    const uint64_t p = mine[0], h = mine[1], b = mine[2];
    NC(g_nccl.Send(t->pair_nhits.p, 4 * p, ncclUint8, 0, cs->comm, st));
    NC(g_nccl.Send(t->pair_fp.p, 16 * p, ncclUint8, 0, cs->comm, st));
    NC(g_nccl.Send(t->pair_hit_off.p, 8 * p, ncclUint8, 0, cs->comm, st));
    NC(g_nccl.Send(t->pair_name_off.p + 1, 8 * p, ncclUint8, 0, cs->comm, st));
    if (h) NC(g_nccl.Send(t->hits.p, 8 * h, ncclUint8, 0, cs->comm, st));
    if (b) NC(g_nccl.Send(t->name_blob.p, b, ncclUint8, 0, cs->comm, st));
  }
  NC(g_nccl.GroupEnd());
  CC(cudaStreamSynchronize(st));
  if (R == 0) {
    for (int r = 1; r < W; ++r) {
      const uint64_t *info = all.data() + (size_t)INFO_WORDS * r;
      if ((rc = tail_absorb_commit(t, info[0], info[1], info[2], st, launches))) return ctx_fail(rc, "tail: %s", tail_error());
    }
  } else {
    tail_reset(t);                                              // this rank's pairs now live on rank 0
  }
  return 0;
}
// rank 0 holds every pair: it finishes alone, the others contribute zeros to the allreduce that hands everyone the counts
int finish_gathered(smash_ctx *c, CommState *cs, TailState *t, cudaStream_t st, uint64_t *launches, int64_t *counts, void *counts_device,
                    smash_tail_stats *st_out) {
  int rc;
  const uint64_t nb = ctx_n_bins(c);
  if ((rc = cs->counts.ensure(nb + 8))) return rc;
  CC(cudaMemsetAsync(cs->counts.p, 0, 8 * (nb + 8), st));
  uint64_t *h_stats = cs->h_small;
  if (cs->rank == 0) {
    smash_tail_stats stl{};
    if ((rc = tail_finish(t, nullptr, (int64_t *)cs->counts.p, &stl, st, launches))) return ctx_fail(rc, "tail: %s", tail_error());
    h_stats[0] = stl.total_reads; h_stats[1] = stl.dups_removed; h_stats[2] = stl.reads_kept; h_stats[3] = stl.n_dupe_pairs; h_stats[4] = stl.n_non_dupe_pairs; h_stats[5] = stl.n_positions;
    CC(cudaMemcpyAsync(cs->counts.p + nb, h_stats, 48, cudaMemcpyHostToDevice, st));
  }
  NC(g_nccl.AllReduce(cs->counts.p, cs->counts.p, nb + 6, ncclUint64, ncclSum, cs->comm, st));
  if (counts) CC(cudaMemcpyAsync(counts, cs->counts.p, 8 * nb, cudaMemcpyDeviceToHost, st));
  if (counts_device) CC(cudaMemcpyAsync(counts_device, cs->counts.p, 8 * nb, cudaMemcpyDeviceToDevice, st));
  CC(cudaMemcpyAsync(h_stats, cs->counts.p + nb, 48, cudaMemcpyDeviceToHost, st));
  CC(cudaStreamSynchronize(st));
  if (st_out) {
    st_out->total_reads = h_stats[0]; st_out->dups_removed = h_stats[1]; st_out->reads_kept = h_stats[2];
    st_out->n_dupe_pairs = h_stats[3]; st_out->n_non_dupe_pairs = h_stats[4]; st_out->n_positions = h_stats[5];
  }
  return 0;
}

}  // namespace

extern "C" int smash_comm_unique_id(void *id, size_t cap) {
  if (!id || cap < sizeof(ncclUniqueId)) return ctx_fail(SMASH_ERR_ARG, "unique id buffer must hold %zu bytes", sizeof(ncclUniqueId));
  if (int rc = nccl_load()) return rc;
  ncclUniqueId u;
  NC(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return 0;
}
extern "C" int smash_comm_init_rank(smash_ctx *c, int rank, int world, const void *id) {
  if (!c || !id || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world) return ctx_fail(SMASH_ERR_ARG, "bad argument");
  if (*ctx_comm_slot(c)) return ctx_fail(SMASH_ERR_STATE, "context already has a communicator");
  if (int rc = nccl_load()) return rc;
  CC(cudaSetDevice(ctx_device(c)));
  ncclUniqueId u; memcpy(&u, id, sizeof u);
  ncclComm_t comm = nullptr;
  NC(g_nccl.CommInitRank(&comm, world, u, rank));
  return comm_attach(c, comm, rank, world);
}
extern "C" int smash_comm_init_all(smash_ctx *const *ctxs, int n) {
  if (!ctxs || n < 1 || n > MAX_WORLD) return ctx_fail(SMASH_ERR_ARG, "bad argument");
  if (int rc = nccl_load()) return rc;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) { if (!ctxs[i] || *ctx_comm_slot(ctxs[i])) return ctx_fail(SMASH_ERR_STATE, "context %d missing or already in a communicator", i); devs[i] = ctx_device(ctxs[i]); }
  std::vector<ncclComm_t> comms(n, nullptr);
  NC(g_nccl.CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) if (int rc = comm_attach(ctxs[i], comms[i], i, n)) return rc;
  return 0;
}
extern "C" int smash_comm_destroy(smash_ctx *c) {
  if (!c) return 0;
  CommState *cs = (CommState *)*ctx_comm_slot(c);
  if (!cs) return 0;
  cudaSetDevice(ctx_device(c));
  if (cs->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(cs->comm);
  cs->send.release(); cs->recv.release(); cs->fp1.release(); cs->fp1_sorted.release(); cs->mins.release(); cs->back.release();
  cs->verdict.release(); cs->small.release(); cs->counts.release(); cs->info.release(); cs->perm.release(); cs->idx.release(); cs->idx_sorted.release(); cs->sort_tmp.release();
  if (cs->h_small) cudaFreeHost(cs->h_small);
  delete cs;
  *ctx_comm_slot(c) = nullptr;
  return 0;
}
extern "C" int smash_comm_rank(const smash_ctx *c, int *rank, int *world) {
  CommState *cs = c ? (CommState *)*ctx_comm_slot(const_cast<smash_ctx *>(c)) : nullptr;
  if (rank) *rank = cs ? cs->rank : 0;
  if (world) *world = cs ? cs->world : 1;
  return 0;
}

// Collective: every rank of the communicator calls it once all its batches are collected.
extern "C" int smash_bins_finish(smash_ctx *c, uint64_t ordinal_base, int64_t *counts, void *counts_device, smash_tail_stats *st_out) {
  if (!c) return ctx_fail(SMASH_ERR_ARG, "null argument");
  CommState *cs = (CommState *)*ctx_comm_slot(c);
  if (!cs || cs->world == 1) {                                  // no communicator: the single-GPU tail
    if (int rc = smash_tail_finish(c, counts, counts_device, st_out)) return rc;
    return 0;
  }
  if (int rc = ctx_require_idle(c)) return rc;
  CC(cudaSetDevice(ctx_device(c)));
  TailState *t = ctx_tail(c);
  cudaStream_t st = ctx_stream(c);
  uint64_t *launches = ctx_launches(c);
  CC(cudaDeviceSynchronize());
  const int W = cs->world, R = cs->rank;
  int rc;
  // 0. are the shards, one after the other, a run of the name order (smash_mapping.sh:23)?  If not, rank 0 finishes alone.
  {
    bool gathered = false;
    if ((rc = gather_if_unsorted(cs, t, st, launches, &gathered))) return rc;
    if (gathered) return finish_gathered(c, cs, t, st, launches, counts, counts_device, st_out);
  }
  // 1. this rank's dupe-set keys, bucketed by owner
  const uint64_t *keys = nullptr; uint64_t n = 0;
  if ((rc = tail_export_keys(t, ordinal_base, &keys, &n, st, launches))) return ctx_fail(rc, "tail: %s", tail_error());
  if (n >= 0xffffffffull) return ctx_fail(SMASH_ERR_ARG, "too many dupe-set pairs on one rank");
  if ((rc = cs->small.ensure(4 * MAX_WORLD + (size_t)W * W + 64)) || (rc = cs->send.ensure(3 * n + 3)) || (rc = cs->perm.ensure(n + 1))) return rc;
  uint64_t *d_cnt = cs->small.p, *d_off = cs->small.p + MAX_WORLD, *d_cur = cs->small.p + 2 * MAX_WORLD, *d_mat = cs->small.p + 4 * MAX_WORLD;
  CC(cudaMemsetAsync(cs->small.p, 0, 8 * 4 * MAX_WORLD, st));
  if (n) { k_owner_count<<<grid_for(n), 256, 0, st>>>(keys, n, W, d_cnt); ++*launches; }
  NC(g_nccl.AllGather(d_cnt, d_mat, (size_t)W, ncclUint64, cs->comm, st));      // mat[src][dst]
  CC(cudaMemcpyAsync(cs->h_small, d_mat, 8 * (size_t)W * W, cudaMemcpyDeviceToHost, st));
  CC(cudaStreamSynchronize(st));
  std::vector<uint64_t> send_cnt(W), send_off(W), recv_cnt(W), recv_off(W);
  uint64_t n_recv = 0, acc = 0;
  for (int r = 0; r < W; ++r) { send_cnt[r] = cs->h_small[(size_t)R * W + r]; send_off[r] = acc; acc += send_cnt[r]; }
  for (int r = 0; r < W; ++r) { recv_cnt[r] = cs->h_small[(size_t)r * W + R]; recv_off[r] = n_recv; n_recv += recv_cnt[r]; }
  if (acc != n) return ctx_fail(SMASH_ERR_STATE, "internal: bucket counts do not add up");
  if (n_recv >= 0xffffffffull) return ctx_fail(SMASH_ERR_ARG, "too many keys for one owner");
  CC(cudaMemcpyAsync(d_off, send_off.data(), 8 * (size_t)W, cudaMemcpyHostToDevice, st));
  if (n) { k_owner_scatter<<<grid_for(n), 256, 0, st>>>(keys, n, W, d_off, d_cur, cs->send.p, cs->perm.p); ++*launches; }
  // 2. buckets to their owners; 3. the owner resolves first-wins per (fp1, fp2); 4. verdicts back
  if ((rc = cs->recv.ensure(3 * n_recv + 3)) || (rc = cs->fp1.ensure(n_recv + 1)) || (rc = cs->fp1_sorted.ensure(n_recv + 1)) ||
      (rc = cs->idx.ensure(n_recv + 1)) || (rc = cs->idx_sorted.ensure(n_recv + 1)) || (rc = cs->mins.ensure(n_recv + 1)) ||
      (rc = cs->back.ensure(n + 1)) || (rc = cs->verdict.ensure(n + 1)))
    return rc;
  if ((rc = all_to_all(cs, cs->send.p, send_cnt.data(), send_off.data(), cs->recv.p, recv_cnt.data(), recv_off.data(), 3, st))) return rc;
  if (n_recv) {
    k_take_fp1<<<grid_for(n_recv), 256, 0, st>>>(cs->recv.p, n_recv, cs->fp1.p, cs->idx.p);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, cs->fp1.p, cs->fp1_sorted.p, cs->idx.p, cs->idx_sorted.p, (int64_t)n_recv, 0, 64, st);
    if ((rc = cs->sort_tmp.ensure(tmp + 16))) return rc;
    CC(cub::DeviceRadixSort::SortPairs(cs->sort_tmp.p, tmp, cs->fp1.p, cs->fp1_sorted.p, cs->idx.p, cs->idx_sorted.p, (int64_t)n_recv, 0, 64, st));
    k_group_min<<<grid_for(n_recv), 256, 0, st>>>(cs->recv.p, cs->fp1_sorted.p, cs->idx_sorted.p, n_recv, cs->mins.p);
    *launches += 4;
  }
  if ((rc = all_to_all(cs, cs->mins.p, recv_cnt.data(), recv_off.data(), cs->back.p, send_cnt.data(), send_off.data(), 1, st))) return rc;
  if (n) { k_unpermute<<<grid_for(n), 256, 0, st>>>(cs->back.p, cs->perm.p, n, cs->verdict.p); ++*launches; }
  CC(cudaStreamSynchronize(st));
  // 5. local duplicate removal + ordered compaction with the global verdicts
  smash_tail_edge edge{};
  static const uint64_t dummy = 0;
  if ((rc = tail_phase_a(t, ordinal_base, nullptr, 0, &edge, st, launches, n ? cs->verdict.p : &dummy, true))) return ctx_fail(rc, "tail: %s", tail_error());
  // 6. shard edges: the last filtered position of the nearest lower rank that has one (varbin.py:56-58)
  uint64_t *h_edge = cs->h_small, *h_edges = cs->h_small + 8;
  h_edge[0] = edge.n_filtered; h_edge[1] = (uint64_t)edge.first_pos; h_edge[2] = (uint64_t)edge.last_pos;
  uint64_t *d_edge = cs->small.p, *d_edges = cs->small.p + 8;
  CC(cudaMemcpyAsync(d_edge, h_edge, 24, cudaMemcpyHostToDevice, st));
  NC(g_nccl.AllGather(d_edge, d_edges, 3, ncclUint64, cs->comm, st));
  CC(cudaMemcpyAsync(h_edges, d_edges, 24 * (size_t)W, cudaMemcpyDeviceToHost, st));
  CC(cudaStreamSynchronize(st));
  int has_prev = 0; int64_t prev = 0;
  for (int r = R - 1; r >= 0; --r) if (h_edges[3 * r]) { has_prev = 1; prev = (int64_t)h_edges[3 * r + 2]; break; }
  // 7. counts of this shard, then the ONE allreduce (+ the stats counters)
  const uint64_t nb = ctx_n_bins(c);
  if ((rc = cs->counts.ensure(nb + 8))) return rc;
  smash_tail_stats stl{};
  if ((rc = tail_phase_b(t, has_prev, prev, nullptr, (int64_t *)cs->counts.p, &stl, st, launches))) return ctx_fail(rc, "tail: %s", tail_error());
  uint64_t *h_stats = cs->h_small;
  h_stats[0] = stl.total_reads; h_stats[1] = stl.dups_removed; h_stats[2] = stl.reads_kept; h_stats[3] = stl.n_dupe_pairs; h_stats[4] = stl.n_non_dupe_pairs; h_stats[5] = stl.n_positions;
  CC(cudaMemcpyAsync(cs->counts.p + nb, h_stats, 48, cudaMemcpyHostToDevice, st));
  NC(g_nccl.AllReduce(cs->counts.p, cs->counts.p, nb + 6, ncclUint64, ncclSum, cs->comm, st));
  if (counts) CC(cudaMemcpyAsync(counts, cs->counts.p, 8 * nb, cudaMemcpyDeviceToHost, st));
  if (counts_device) CC(cudaMemcpyAsync(counts_device, cs->counts.p, 8 * nb, cudaMemcpyDeviceToDevice, st));
  CC(cudaMemcpyAsync(h_stats, cs->counts.p + nb, 48, cudaMemcpyDeviceToHost, st));
  CC(cudaStreamSynchronize(st));
  if (st_out) {
    st_out->total_reads = h_stats[0]; st_out->dups_removed = h_stats[1]; st_out->reads_kept = h_stats[2];
    st_out->n_dupe_pairs = h_stats[3]; st_out->n_non_dupe_pairs = h_stats[4]; st_out->n_positions = h_stats[5];
  }
  return 0;
}
