// gcnorm.cu -- GC normalisation of the bin counts on the GPU: the head of cbs.segment01 (cbs.r:18-25) with lowess.gc
// (cbs.r:3-7), SURVEY.md §8(f)4.
//
//     a        <- bincount + 1
//     ratio    <- a / mean(a[autosomes])
//     lowratio <- exp(log(ratio) - approx(lowess(gc, log(ratio), f = 0.05), xout = gc)$y)
//
// `stats::lowess` is R's C translation (clowess / lowest) of Cleveland's LOWESS with iter = 3 and delta = 1 % of the x
// range; `stats::approx` interpolates linearly with tied x averaged.  What depends only on x = gc.content (a property of
// the bin file, fixed for a run) is worked out once on the host when the object is created:
//   * the sort by x, and clowess's walk over the sorted points: which points get a local fit (the others lie within delta
//     of the last fit and are interpolated, or repeat its x and copy it), and each fit's window [nleft, nright];
//   * approx's groups of tied x.
// Per call the GPU does the arithmetic: ratio and log, then per robustness iteration one thread per fitted point running
// `lowest` with its sums in the reference's order (so the result does not depend on the launch geometry), the
// interpolation of the skipped points, the residuals, the median of their absolute values (device radix sort) and the
// bisquare weights; finally the tie averages and exp().  A few hundred fits of a few thousand points each: the stage is
// small next to the mapping, it is here so that the counts never have to leave the device before they are normalised.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#if !defined(SMASH_CUDA_SHIM)
#include <cub/device/device_radix_sort.cuh>
#endif

#include "../../include/smash_b200.h"
#include "ctx_internal.h"

namespace {

#define GCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(SMASH_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

struct Fit { uint32_t i, nleft, nright, pad; };                // a fitted point (sorted index) and its window (inclusive)
struct Fill { uint32_t ka, kb; double alpha; };                // ys[j] = alpha * fit[kb] + (1 - alpha) * fit[ka]  (ka == kb: copy)

}  // namespace

struct smash_gcnorm {
  int device = 0;
  uint64_t n = 0, n_fit = 0, n_grp = 0, n_auto = 0;
  int nsteps = 3;
  cudaStream_t st = nullptr;
  // x-dependent plan
  double *x = nullptr;             // sorted gc.content
  uint32_t *order = nullptr;       // sorted position -> bin
  uint8_t *autosome = nullptr;     // per bin
  Fit *fits = nullptr;
  Fill *fill = nullptr;            // per sorted point
  uint32_t *grp_of = nullptr;      // per sorted point: tie group
  uint32_t *grp_start = nullptr;   // n_grp + 1
  // per call
  int64_t *counts = nullptr;
  double *ratio = nullptr, *y = nullptr, *ys = nullptr, *yfit = nullptr, *rw = nullptr, *res = nullptr, *absres = nullptr, *sorted = nullptr,
         *uy = nullptr, *low = nullptr, *scal = nullptr;   // scal: [0] autosome sum, [1] sum |res|, [2] stop flag
  void *sort_tmp = nullptr; size_t sort_tmp_bytes = 0;
  double *h_out = nullptr;         // pinned, 2 n
};

namespace {

__global__ void k_gc_sum(const int64_t *__restrict__ counts, const uint8_t *__restrict__ autosome, uint64_t n, unsigned long long *sum) {
  unsigned long long loc = 0;
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (uint64_t)gridDim.x * blockDim.x)
    if (autosome[b]) loc += (unsigned long long)(counts[b] + 1);
  for (int o = 16; o; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
  if ((threadIdx.x & 31) == 0 && loc) atomicAdd(sum, loc);       // integers: exact in any order
}
// ratio per bin, y = log(ratio) in sorted order, weights reset
__global__ void k_gc_ratio(const int64_t *__restrict__ counts, const uint32_t *__restrict__ order, uint64_t n, const unsigned long long *sum,
                           uint64_t n_auto, double *__restrict__ ratio, double *__restrict__ y, double *__restrict__ scal) {
  const double mean = (double)*sum / (double)n_auto;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = order[j];
    const double r = (double)(counts[b] + 1) / mean;
    ratio[b] = r;
    y[j] = log(r);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { scal[1] = 0.0; scal[2] = 0.0; }
}
// `lowest` (lowess.c) for one fitted point; the weights are recomputed in every pass instead of being kept in an array
__device__ __forceinline__ double gc_weight(double xj, double xs, double h, double h9, double h1, bool userw, const double *rw, uint64_t j, bool *in) {
  const double r = fabs(xj - xs);
  *in = r <= h9;
  if (!*in) return 0.0;
  double w;
  if (r <= h1) w = 1.0;
  else { const double t = r / h; const double u = 1.0 - t * t * t; w = u * u * u; }
  if (userw) w *= rw[j];
  return w;
}
__global__ void k_gc_fit(const double *__restrict__ x, const double *__restrict__ y, uint64_t n, const Fit *__restrict__ fits, uint64_t n_fit,
                         const double *__restrict__ rw, int userw, const double *__restrict__ scal, double *__restrict__ yfit) {
  if (scal[2] != 0.0) return;                                   // the robustness loop has ended (cmad ~ 0): keep the last fit
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_fit) return;
  const Fit f = fits[k];
  const double xs = x[f.i];
  const double range = x[n - 1] - x[0];
  const double h = fmax(xs - x[f.nleft], x[f.nright] - xs);
  const double h9 = 0.999 * h, h1 = 0.001 * h;
  // pass 1: sum of weights, and where the window ends (ties of x[nright] beyond it are included, as in the reference)
  double a = 0.0;
  uint64_t j = f.nleft;
  while (j < n) {
    bool in;
    const double wj = gc_weight(x[j], xs, h, h9, h1, userw != 0, rw, j, &in);
    if (in) a += wj;
    else if (x[j] > xs) break;
    ++j;
  }
  const uint64_t nrt = j;                                       // one past the last point looked at
  if (a <= 0.0) { yfit[k] = y[f.i]; return; }
  double slope_b = 0.0, centre = 0.0;
  bool linear = false;
  if (h > 0.0) {
    double sx = 0.0;
    for (j = f.nleft; j < nrt; ++j) { bool in; const double wj = gc_weight(x[j], xs, h, h9, h1, userw != 0, rw, j, &in) / a; sx += wj * x[j]; }
    centre = sx;
    double b = xs - centre, c = 0.0;
    for (j = f.nleft; j < nrt; ++j) { bool in; const double wj = gc_weight(x[j], xs, h, h9, h1, userw != 0, rw, j, &in) / a; c += wj * (x[j] - centre) * (x[j] - centre); }
    if (sqrt(c) > 0.001 * range) { slope_b = b / c; linear = true; }
  }
  double ysum = 0.0;
  for (j = f.nleft; j < nrt; ++j) {
    bool in;
    double wj = gc_weight(x[j], xs, h, h9, h1, userw != 0, rw, j, &in) / a;
    if (linear) wj *= (slope_b * (x[j] - centre) + 1.0);
    ysum += wj * y[j];
  }
  yfit[k] = ysum;
}
// every sorted point from the fitted ones; residuals
__global__ void k_gc_fill(const Fill *__restrict__ fill, const double *__restrict__ yfit, const double *__restrict__ y, uint64_t n,
                          const double *__restrict__ scal, double *__restrict__ ys, double *__restrict__ res, double *__restrict__ absres) {
  if (scal[2] != 0.0) return;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    const Fill f = fill[j];
    const double v = f.ka == f.kb ? yfit[f.ka] : f.alpha * yfit[f.kb] + (1.0 - f.alpha) * yfit[f.ka];
    ys[j] = v;
    const double r = y[j] - v;
    res[j] = r; absres[j] = fabs(r);
  }
}
// sum |res| in a fixed order (one block, strided partial sums then a tree)
__global__ void k_gc_abs_sum(const double *__restrict__ absres, uint64_t n, double *__restrict__ scal) {
  if (scal[2] != 0.0) return;
  __shared__ double part[1024];
  double loc = 0.0;
  for (uint64_t j = threadIdx.x; j < n; j += blockDim.x) loc += absres[j];
  part[threadIdx.x] = loc;
  __syncthreads();
  for (int o = blockDim.x / 2; o; o >>= 1) { if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) scal[1] = part[0];
}
// robustness weights from the median absolute residual (sorted: |res| ascending)
__global__ void k_gc_weights(const double *__restrict__ res, const double *__restrict__ sorted, uint64_t n, double *__restrict__ scal, double *__restrict__ rw) {
  if (scal[2] != 0.0) return;
  const uint64_t m1 = n / 2;
  const double cmad = (n % 2 == 0) ? 3.0 * (sorted[m1] + sorted[n - m1 - 1]) : 6.0 * sorted[m1];
  const double sc = scal[1] / (double)n;
  const bool stop = cmad < 1e-7 * sc;
  const double c9 = 0.999 * cmad, c1 = 0.001 * cmad;
  if (!stop)
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
      const double r = fabs(res[j]);
      double wv;
      if (r <= c1) wv = 1.0;
      else if (r <= c9) { const double t = r / cmad; const double u = 1.0 - t * t; wv = u * u; }
      else wv = 0.0;
      rw[j] = wv;
    }
  // every thread has read scal[2] == 0 before any thread sets it: the flag is written by a separate one-thread kernel
}
__global__ void k_gc_stop(const double *__restrict__ sorted, uint64_t n, double *__restrict__ scal) {
  if (scal[2] != 0.0) return;
  const uint64_t m1 = n / 2;
  const double cmad = (n % 2 == 0) ? 3.0 * (sorted[m1] + sorted[n - m1 - 1]) : 6.0 * sorted[m1];
  if (cmad < 1e-7 * (scal[1] / (double)n)) scal[2] = 1.0;
}
// approx(): tied x share the mean of their fitted values; lowratio = exp(y - z)
__global__ void k_gc_groups(const double *__restrict__ ys, const uint32_t *__restrict__ grp_start, uint64_t n_grp, double *__restrict__ uy) {
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_grp; g += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t s = grp_start[g], e = grp_start[g + 1];
    double sum = 0.0;
    for (uint32_t j = s; j < e; ++j) sum += ys[j];
    uy[g] = sum / (double)(e - s);
  }
}
__global__ void k_gc_out(const double *__restrict__ y, const double *__restrict__ uy, const uint32_t *__restrict__ grp_of, const uint32_t *__restrict__ order,
                         uint64_t n, double *__restrict__ low) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
    low[order[j]] = exp(y[j] - uy[grp_of[j]]);
}

template <class T> int dalloc(T **p, size_t n) {
  if (cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return ctx_fail(SMASH_ERR_NOMEM, "cudaMalloc(%zu bytes)", n * sizeof(T));
  return 0;
}
template <class T> int upload(T **p, const std::vector<T> &v) {
  if (int rc = dalloc(p, v.size())) return rc;
  if (!v.empty() && cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return ctx_fail(SMASH_ERR_CUDA, "cudaMemcpy");
  return 0;
}
int grid_for(uint64_t n) { const uint64_t g = (n + 255) / 256; return (int)(g < 1 ? 1 : g > 148 * 8 ? 148 * 8 : g); }

}  // namespace

extern "C" int smash_gcnorm_create(int device, const double *gc_content, const uint8_t *autosome, uint64_t n_bins, double f, int iter,
                                   smash_gcnorm **out) {
  if (!gc_content || !autosome || !out || n_bins < 2 || n_bins >= 0xffffffffull || !(f > 0.0) || iter < 0)
    return ctx_fail(SMASH_ERR_ARG, "smash_gcnorm_create: bad argument");
  if (smash_device_count() <= 0) return ctx_fail(SMASH_ERR_CUDA, "no sm_100 CUDA device available (this library has no CPU fallback)");
  GCU(cudaSetDevice(device));
  const uint64_t n = n_bins;
  for (uint64_t b = 0; b < n; ++b) if (!(gc_content[b] == gc_content[b])) return ctx_fail(SMASH_ERR_DATA, "gc.content of bin %llu is not a number", (unsigned long long)b);
  // order(x): stable
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return gc_content[a] < gc_content[b]; });
  std::vector<double> x(n);
  for (uint64_t j = 0; j < n; ++j) x[j] = gc_content[order[j]];
  const double delta = 0.01 * (x[n - 1] - x[0]);                       // lowess(): delta = 0.01 * diff(range(x))
  int64_t ns = (int64_t)(f * (double)n + 1e-7);
  ns = std::max<int64_t>(2, std::min<int64_t>((int64_t)n, ns));
  // clowess's walk (lowess.c): it depends on x alone
  std::vector<Fit> fits;
  std::vector<Fill> fill(n);
  {
    int64_t nleft = 0, nright = ns - 1, last = -1, i = 0;
    uint32_t k_last = 0;                                               // fit whose value ys[last] holds
    for (;;) {
      if (nright < (int64_t)n - 1) {
        const double d1 = x[i] - x[nleft], d2 = x[nright + 1] - x[i];
        if (d1 > d2) { ++nleft; ++nright; continue; }
      }
      const uint32_t k = (uint32_t)fits.size();
      fits.push_back(Fit{(uint32_t)i, (uint32_t)nleft, (uint32_t)nright, 0u});
      fill[i] = Fill{k, k, 0.0};
      if (last < i - 1) {
        const double denom = x[i] - x[last];
        for (int64_t j = last + 1; j < i; ++j) fill[j] = Fill{k_last, k, (x[j] - x[last]) / denom};
      }
      last = i; k_last = k;
      const double cut = x[last] + delta;
      for (i = last + 1; i < (int64_t)n; ++i) {
        if (x[i] > cut) break;
        if (x[i] == x[last]) { fill[i] = Fill{k, k, 0.0}; last = i; }
      }
      i = std::max<int64_t>(last + 1, i - 1);
      if (last >= (int64_t)n - 1) break;
    }
  }
  // approx(): groups of tied x
  std::vector<uint32_t> grp_of(n), grp_start;
  for (uint64_t j = 0; j < n; ++j) {
    if (j == 0 || x[j] != x[j - 1]) grp_start.push_back((uint32_t)j);
    grp_of[j] = (uint32_t)grp_start.size() - 1;
  }
  const uint64_t n_grp = grp_start.size();
  grp_start.push_back((uint32_t)n);
  uint64_t n_auto = 0;
  std::vector<uint8_t> au(autosome, autosome + n);
  for (uint64_t b = 0; b < n; ++b) n_auto += au[b] ? 1 : 0;
  if (!n_auto) return ctx_fail(SMASH_ERR_DATA, "no autosome bins: the mean of cbs.r:21 is undefined");

  smash_gcnorm *g = new smash_gcnorm();
  g->device = device; g->n = n; g->n_fit = fits.size(); g->n_grp = n_grp; g->n_auto = n_auto; g->nsteps = iter;
  int rc = 0;
  if (cudaStreamCreate(&g->st) != cudaSuccess) { delete g; return ctx_fail(SMASH_ERR_CUDA, "cudaStreamCreate"); }
  if ((rc = upload(&g->x, x)) || (rc = upload(&g->order, order)) || (rc = upload(&g->autosome, au)) || (rc = upload(&g->fits, fits)) ||
      (rc = upload(&g->fill, fill)) || (rc = upload(&g->grp_of, grp_of)) || (rc = upload(&g->grp_start, grp_start)) ||
      (rc = dalloc(&g->counts, n)) || (rc = dalloc(&g->ratio, n)) || (rc = dalloc(&g->y, n)) || (rc = dalloc(&g->ys, n)) ||
      (rc = dalloc(&g->yfit, fits.size())) || (rc = dalloc(&g->rw, n)) || (rc = dalloc(&g->res, n)) || (rc = dalloc(&g->absres, n)) ||
      (rc = dalloc(&g->sorted, n)) || (rc = dalloc(&g->uy, n_grp)) || (rc = dalloc(&g->low, n)) || (rc = dalloc(&g->scal, 4))) {
    smash_gcnorm_destroy(g);
    return rc;
  }
#if !defined(SMASH_CUDA_SHIM)
  cub::DeviceRadixSort::SortKeys(nullptr, g->sort_tmp_bytes, g->absres, g->sorted, (int64_t)n, 0, 64, g->st);
#endif
  if (cudaMalloc(&g->sort_tmp, g->sort_tmp_bytes + 16) != cudaSuccess || cudaHostAlloc((void **)&g->h_out, 16 * n, cudaHostAllocDefault) != cudaSuccess) {
    smash_gcnorm_destroy(g);
    return ctx_fail(SMASH_ERR_NOMEM, "smash_gcnorm_create: out of memory");
  }
  *out = g;
  return 0;
}

extern "C" int smash_gcnorm_run(smash_gcnorm *g, const int64_t *counts, const void *counts_device, double *ratio, double *lowratio) {
  if (!g || (!counts && !counts_device)) return ctx_fail(SMASH_ERR_ARG, "smash_gcnorm_run: null argument");
  GCU(cudaSetDevice(g->device));
  cudaStream_t st = g->st;
  const uint64_t n = g->n;
  const int64_t *cnt = (const int64_t *)counts_device;
  if (!cnt) { GCU(cudaMemcpyAsync(g->counts, counts, 8 * n, cudaMemcpyHostToDevice, st)); cnt = g->counts; }
  unsigned long long *isum = reinterpret_cast<unsigned long long *>(g->scal + 3);
  GCU(cudaMemsetAsync(g->scal, 0, 32, st));
  k_gc_sum<<<grid_for(n), 256, 0, st>>>(cnt, g->autosome, n, isum);
  k_gc_ratio<<<grid_for(n), 256, 0, st>>>(cnt, g->order, n, isum, g->n_auto, g->ratio, g->y, g->scal);
  for (int it = 1; it <= g->nsteps + 1; ++it) {
    k_gc_fit<<<(unsigned)((g->n_fit + 63) / 64), 64, 0, st>>>(g->x, g->y, n, g->fits, g->n_fit, g->rw, it > 1 ? 1 : 0, g->scal, g->yfit);
    k_gc_fill<<<grid_for(n), 256, 0, st>>>(g->fill, g->yfit, g->y, n, g->scal, g->ys, g->res, g->absres);
    if (it > g->nsteps) break;
    k_gc_abs_sum<<<1, 1024, 0, st>>>(g->absres, n, g->scal);
#if !defined(SMASH_CUDA_SHIM)
    GCU(cub::DeviceRadixSort::SortKeys(g->sort_tmp, g->sort_tmp_bytes, g->absres, g->sorted, (int64_t)n, 0, 64, st));
#else
    GCU(cudaStreamSynchronize(st));                              // host emulation (tests/emul): device memory is host memory there
    std::copy(g->absres, g->absres + n, g->sorted); std::sort(g->sorted, g->sorted + n);
#endif
    k_gc_weights<<<grid_for(n), 256, 0, st>>>(g->res, g->sorted, n, g->scal, g->rw);
    k_gc_stop<<<1, 1, 0, st>>>(g->sorted, n, g->scal);
  }
  k_gc_groups<<<grid_for(g->n_grp), 256, 0, st>>>(g->ys, g->grp_start, g->n_grp, g->uy);
  k_gc_out<<<grid_for(n), 256, 0, st>>>(g->y, g->uy, g->grp_of, g->order, n, g->low);
  GCU(cudaMemcpyAsync(g->h_out, g->ratio, 8 * n, cudaMemcpyDeviceToHost, st));
  GCU(cudaMemcpyAsync(g->h_out + n, g->low, 8 * n, cudaMemcpyDeviceToHost, st));
  GCU(cudaStreamSynchronize(st));
  GCU(cudaGetLastError());
  if (ratio) memcpy(ratio, g->h_out, 8 * n);
  if (lowratio) memcpy(lowratio, g->h_out + n, 8 * n);
  return 0;
}

extern "C" void smash_gcnorm_destroy(smash_gcnorm *g) {
  if (!g) return;
  cudaSetDevice(g->device);
  void *p[] = {g->x, g->order, g->autosome, g->fits, g->fill, g->grp_of, g->grp_start, g->counts, g->ratio, g->y, g->ys, g->yfit, g->rw, g->res,
               g->absres, g->sorted, g->uy, g->low, g->scal, g->sort_tmp};
  for (void *q : p) if (q) cudaFree(q);
  if (g->h_out) cudaFreeHost(g->h_out);
  if (g->st) cudaStreamDestroy(g->st);
  delete g;
}
