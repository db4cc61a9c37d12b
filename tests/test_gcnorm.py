"""GC normalisation of the bin counts (cbs.r:18-25 with lowess.gc, cbs.r:3-7).

CPU: oracle/gcnorm.py is pinned on published known answers -- the three result vectors printed in the header of the
netlib LOWESS routine (Cleveland's FORTRAN, which R's lowess.c translates line by line) and the fitted values R's own
documentation example `lowess(cars)` prints.  GPU: smash_gcnorm_* against that oracle on the reference's own
sample_bins/50000 gc.content column (tests/golden/sample_bins_50000_gc.txt.gz) with synthetic counts, at 1e-9 relative
(double precision on both sides; the sums are taken in a different order and libm's log/exp differ in the last place).
"""
import gzip
import os

import numpy as np
import pytest

from oracle import gcnorm as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# netlib lowess.f, header comment: test driver data and the YS it must print
NETLIB_X = np.array([1, 2, 3, 4, 5] + [6] * 10 + [8, 10, 12, 14, 50], dtype=np.float64)
NETLIB_Y = np.array([18, 2, 15, 6, 10, 4, 16, 11, 7, 3, 14, 17, 20, 12, 9, 13, 1, 8, 5, 19], dtype=np.float64)
NETLIB_YS = {
    (0.25, 0, 0.0): [13.659, 11.145, 8.701, 9.722, 10.000] + [11.300] * 10 + [13.000, 6.440, 5.596, 5.456, 18.998],
    (0.25, 0, 3.0): [13.659, 12.347, 11.034, 9.722, 10.511] + [11.300] * 10 + [13.000, 6.440, 5.596, 5.456, 18.998],
    (0.25, 2, 0.0): [14.811, 12.115, 8.984, 9.676, 10.000] + [11.346] * 10 + [13.000, 6.734, 5.744, 5.415, 18.998],
}
# R: datasets::cars and what `lowess(cars)` prints (R documentation, example(lowess))
CARS_SPEED = [4, 4, 7, 7, 8, 9, 10, 10, 10, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13, 14, 14, 14, 14, 15, 15, 15, 16, 16, 17, 17, 17, 18, 18,
              18, 18, 19, 19, 19, 20, 20, 20, 20, 20, 22, 23, 24, 24, 24, 24, 25]
CARS_DIST = [2, 10, 4, 22, 16, 10, 18, 26, 34, 17, 28, 14, 20, 24, 28, 26, 34, 34, 46, 26, 36, 60, 80, 20, 26, 54, 32, 40, 32, 40, 50, 42, 56,
             76, 84, 36, 46, 68, 32, 48, 52, 56, 64, 66, 54, 70, 92, 93, 120, 85]
CARS_LOWESS_Y = [4.965459, 4.965459, 13.124495, 13.124495, 15.858633, 18.579691, 21.280313, 21.280313, 21.280313, 24.129277, 24.129277,
                 27.119549, 27.119549, 27.119549, 27.119549, 30.027276, 30.027276, 30.027276, 30.027276, 32.962506, 32.962506, 32.962506,
                 32.962506, 36.757728, 36.757728, 36.757728, 40.435075, 40.435075, 43.463492, 43.463492, 43.463492, 46.885479, 46.885479,
                 46.885479, 46.885479, 50.793152, 50.793152, 50.793152, 56.491224, 56.491224, 56.491224, 56.491224, 56.491224, 67.585824,
                 73.079695, 78.643164, 78.643164, 78.643164, 78.643164, 84.328698]


@pytest.mark.parametrize("setting", sorted(NETLIB_YS))
def test_oracle_lowess_reproduces_the_netlib_test_vectors(setting):
    f, nsteps, delta = setting
    ys = G.clowess(NETLIB_X, NETLIB_Y, f, nsteps, delta)
    assert np.allclose(ys, NETLIB_YS[setting], rtol=0, atol=5.1e-4)          # printed with three decimals


def test_oracle_lowess_reproduces_r_example_cars():
    lx, ly = G.r_lowess(CARS_SPEED, CARS_DIST)                               # R defaults: f = 2/3, iter = 3, delta = 1 % of the range
    assert np.array_equal(lx, np.sort(np.array(CARS_SPEED, dtype=np.float64)))
    assert np.allclose(ly, CARS_LOWESS_Y, rtol=0, atol=5.1e-7)               # printed with six decimals


def test_oracle_approx_averages_ties():
    z = G.r_approx([1.0, 2.0, 2.0, 3.0], [10.0, 20.0, 40.0, 50.0], [1.0, 1.5, 2.0, 3.0])
    assert np.allclose(z, [10.0, 20.0, 30.0, 50.0])


def _gc_table():
    rows = [l.split("\t") for l in gzip.open(os.path.join(GOLDEN, "sample_bins_50000_gc.txt.gz"), "rt").read().splitlines()[1:]]
    return [r[0] for r in rows], np.array([float(r[1]) for r in rows])


def _counts(gc, chroms, seed, scale):
    """bin counts with a GC bias, copy-number steps, noise, a few empty and a few huge bins"""
    rng = np.random.default_rng(seed)
    bias = np.exp(-((gc - 0.60) / 0.15) ** 2)
    cn = np.where(np.arange(len(gc)) % 7000 < 900, 1.5, 1.0)
    c = rng.poisson(scale * bias * cn).astype(np.int64)
    c[rng.integers(0, len(gc), 40)] = 0
    c[rng.integers(0, len(gc), 5)] *= 50
    return c


def test_oracle_gc_normalise_flattens_a_gc_bias():
    chroms, gc = _gc_table()
    counts = _counts(gc, chroms, 1, 400.0)
    ratio, low = G.gc_normalise(counts, gc, chroms)
    mid = (gc > 0.36) & (gc < 0.55) & (np.arange(len(gc)) % 7000 >= 900)
    lo_gc, hi_gc = mid & (gc < 0.40), mid & (gc > 0.50)
    assert abs(np.median(ratio[lo_gc]) / np.median(ratio[hi_gc]) - 1.0) > 0.1      # the bias is there ..
    assert abs(np.median(low[lo_gc]) / np.median(low[hi_gc]) - 1.0) < 0.02         # .. and gone after the lowess step


@pytest.mark.gpu
@pytest.mark.parametrize("seed,scale", [(1, 400.0), (2, 3.0), (3, 40000.0)])
def test_gpu_gc_normalise_equals_oracle(seed, scale):
    from smash_paper_b200 import api
    chroms, gc = _gc_table()
    counts = _counts(gc, chroms, seed, scale)
    g = api.GcNorm(gc, chroms)
    try:
        ratio, low = g.run(counts)
        oratio, olow = G.gc_normalise(counts, gc, chroms)
        assert np.allclose(ratio, oratio, rtol=1e-14, atol=0)
        assert np.allclose(low, olow, rtol=1e-9, atol=0)
        ratio2, low2 = g.run(counts)                                               # same object, same answer
        assert np.array_equal(low, low2) and np.array_equal(ratio, ratio2)
    finally:
        g.close()


@pytest.mark.gpu
def test_gpu_gc_normalise_small_tied_and_device_counts():
    """few bins, heavy ties in gc (tie copies + approx's averaging), all-equal counts (the robustness loop ends early:
    residuals ~ 0), and counts handed over as a device pointer."""
    import torch
    from smash_paper_b200 import api
    rng = np.random.default_rng(9)
    gc = np.round(rng.uniform(0.3, 0.6, 600), 2)
    chroms = ["chr%d" % (1 + i % 22) if i % 25 else "chrX" for i in range(600)]
    for counts in (rng.poisson(200, 600).astype(np.int64), np.full(600, 77, dtype=np.int64)):
        g = api.GcNorm(gc, chroms, f=0.3)
        try:
            dev = torch.from_numpy(counts).cuda()
            ratio, low = g.run(counts_device_ptr=dev.data_ptr())
            lx, ly = G.r_lowess(gc, np.log((counts + 1.0) / np.mean((counts + 1.0)[np.array([c != "chrX" for c in chroms])])), f=0.3)
            z = G.r_approx(lx, ly, gc)
            olow = np.exp(np.log(ratio) - z)
            assert np.allclose(low, olow, rtol=1e-9, atol=0)
        finally:
            g.close()
