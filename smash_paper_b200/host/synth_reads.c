/* synth_reads.c -- fast deterministic generator of chimeric SMASH-like read pairs (bench input).
 * Same recipe as synth.make_reads: every read = 3..8 fragments from uniform random
 * loci/strands of the forward genome, substitutions, 'Z' bases (fastqs_to_sam.cpp:69 N->Z), a few
 * fully random reads, a few exact duplicate pairs.  Every read is a pure function of
 * (seed, global pair index), so the output does not depend on the thread count. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL; return x ^ (x >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t next(rng_t *r) { r->s += 0x9e3779b97f4a7c15ULL; return mix(r->s); }
static inline uint8_t comp(uint8_t c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

typedef struct {
  const uint8_t *genome; uint64_t G; uint64_t seed, first_pair, n_pairs; int q, fmin, fmax;
  uint32_t sub_thr, z_thr, rnd_thr, dup_thr;   /* thresholds on 24-bit randoms */
  uint8_t *seq, *qual; uint64_t lo, hi;         /* pair range of this thread */
} job_t;

static void gen_read(const job_t *J, uint64_t pair, int mate, uint8_t *out, uint8_t *qual) {
  rng_t r = { mix(J->seed * 0x100000001b3ULL + 2 * pair + (uint64_t)mate) };
  const int q = J->q;
  static const char ACGT[4] = {'A', 'C', 'G', 'T'};
  if ((next(&r) & 0xffffff) < J->rnd_thr) {
    for (int j = 0; j < q; ++j) out[j] = (uint8_t)ACGT[next(&r) & 3];
  } else {
    int nf = J->fmin + (int)(next(&r) % (uint64_t)(J->fmax - J->fmin + 1));
    int cuts[16]; int nc = nf - 1;
    for (int i = 0; i < nc; ++i) cuts[i] = 1 + (int)(next(&r) % (uint64_t)(q - 1));
    for (int i = 1; i < nc; ++i) { int v = cuts[i], k = i - 1; while (k >= 0 && cuts[k] > v) { cuts[k + 1] = cuts[k]; --k; } cuts[k + 1] = v; }
    int fs = 0;
    for (int f = 0; f <= nc; ++f) {
      int fe = f < nc ? cuts[f] : q;
      if (fe > fs) {
        uint64_t lo = next(&r) % (J->G - (uint64_t)q);
        int rev = (int)(next(&r) & 1);
        for (int j = fs; j < fe; ++j)
          out[j] = rev ? comp(J->genome[lo + (uint64_t)(fe - 1 - j)]) : J->genome[lo + (uint64_t)(j - fs)];
      }
      fs = fe;
    }
  }
  for (int j = 0; j < q; ++j) {
    uint64_t x = next(&r);
    if ((x & 0xffffff) < J->sub_thr) out[j] = (uint8_t)ACGT[(x >> 24) & 3];
    if (((x >> 26) & 0xffffff) < J->z_thr || out[j] == 'N') out[j] = 'Z';
    qual[j] = (uint8_t)(35 + ((x >> 50) % 39));
  }
}

static void *worker(void *arg) {
  job_t *J = (job_t *)arg;
  for (uint64_t p = J->lo; p < J->hi; ++p) {
    uint64_t gp = J->first_pair + p, src = gp;
    /* a few exact duplicate pairs: pair gp copies pair gp - d (d in 1..49) */
    uint64_t h = mix(J->seed ^ (gp * 0x9e3779b97f4a7c15ULL));
    if ((h & 0xffffff) < J->dup_thr) { uint64_t d = 1 + ((h >> 24) % 49); if (gp >= d) src = gp - d; }
    for (int m = 0; m < 2; ++m)
      gen_read(J, src, m, J->seq + (2 * p + (uint64_t)m) * (uint64_t)J->q, J->qual + (2 * p + (uint64_t)m) * (uint64_t)J->q);
  }
  return NULL;
}

void synth_reads(const uint8_t *genome, uint64_t G, uint64_t seed, uint64_t first_pair, uint64_t n_pairs, int q,
                 int fmin, int fmax, double sub_rate, double z_rate, double random_frac, double dup_frac,
                 uint8_t *seq, uint8_t *qual, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  job_t J[64]; pthread_t th[64];
  for (int t = 0; t < n_threads; ++t) {
    J[t] = (job_t){genome, G, seed, first_pair, n_pairs, q, fmin, fmax,
                   (uint32_t)(sub_rate * 16777216.0), (uint32_t)(z_rate * 16777216.0),
                   (uint32_t)(random_frac * 16777216.0), (uint32_t)(dup_frac * 16777216.0), seq, qual,
                   n_pairs * (uint64_t)t / (uint64_t)n_threads, n_pairs * (uint64_t)(t + 1) / (uint64_t)n_threads};
    pthread_create(&th[t], NULL, worker, &J[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
}
