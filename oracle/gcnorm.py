"""oracle/gcnorm.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the GC normalisation at the head of cbs.segment01 (cbs.r:18-25 with lowess.gc, cbs.r:3-7):

    a        <- bincount + 1
    ratio    <- a / mean(a[autosomes])                    # chrom.numeric < 23 (chrX = 23, chrY = 24)
    lowratio <- exp(log(ratio) - approx(lowess(gc, log(ratio), f = 0.05), xout = gc)$y)

R is absent from this image and from /root/reference (cbs.r needs R + DNAcopy), so the arithmetic lives in a
third-party dependency: `stats::lowess` is R's C translation (src/library/stats/src/lowess.c, `clowess`/`lowest`) of
Cleveland's LOWESS (netlib `go/lowess.f`, 1979/1985), called with the R defaults iter = 3, delta = 0.01 * diff(range(x))
after sorting by x; `stats::approx` is linear interpolation with ties averaged (regularize.values, ties = mean).
Both are restated here from the published algorithm.  PINNED on the known-answer vectors printed in the header of the
netlib routine (the 20-point test driver, three settings of F / NSTEPS / DELTA; tests/test_gcnorm.py) -- R's routine is a
line-by-line translation of that FORTRAN.  The cbs.r wrapper around it (three lines of arithmetic) has no reference
output here: floating-point parity of the full stage is checked GPU-vs-this-oracle at 1e-12 relative.
"""
from __future__ import annotations

import numpy as np


def _lowest(x, y, n, xs, nleft, nright, w, userw, rw):
    """`lowest` of lowess.c: fitted value at xs from the points nleft..nright (0-based, inclusive)."""
    rng = x[n - 1] - x[0]
    h = max(xs - x[nleft], x[nright] - xs)
    h9, h1 = 0.999 * h, 0.001 * h
    a = 0.0
    j = nleft
    while j < n:
        w[j] = 0.0
        r = abs(x[j] - xs)
        if r <= h9:
            w[j] = 1.0 if r <= h1 else (1.0 - (r / h) ** 3) ** 3
            if userw:
                w[j] *= rw[j]
            a += w[j]
        elif x[j] > xs:
            break
        j += 1
    nrt = j - 1
    if a <= 0.0:
        return 0.0, False
    w[nleft:nrt + 1] /= a
    if h > 0.0:
        a = float(np.dot(w[nleft:nrt + 1], x[nleft:nrt + 1]))
        b = xs - a
        c = float(np.dot(w[nleft:nrt + 1], (x[nleft:nrt + 1] - a) ** 2))
        if np.sqrt(c) > 0.001 * rng:
            b /= c
            w[nleft:nrt + 1] *= (b * (x[nleft:nrt + 1] - a) + 1.0)
    return float(np.dot(w[nleft:nrt + 1], y[nleft:nrt + 1])), True


def clowess(x, y, f, nsteps, delta):
    """`clowess` of lowess.c on x sorted ascending.  Returns the fitted values."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    ys = np.zeros(n)
    if n < 2:
        ys[:] = y
        return ys
    ns = max(2, min(n, int(f * n + 1e-7)))
    rw = np.zeros(n)
    res = np.zeros(n)
    w = np.zeros(n)
    it = 1
    while it <= nsteps + 1:
        nleft, nright, last, i = 0, ns - 1, -1, 0
        while True:
            if nright < n - 1:
                d1 = x[i] - x[nleft]
                d2 = x[nright + 1] - x[i]
                if d1 > d2:
                    nleft += 1
                    nright += 1
                    continue
            ys[i], ok = _lowest(x, y, n, x[i], nleft, nright, w, it > 1, rw)
            if not ok:
                ys[i] = y[i]
            if last < i - 1:
                denom = x[i] - x[last]
                for j in range(last + 1, i):
                    alpha = (x[j] - x[last]) / denom
                    ys[j] = alpha * ys[i] + (1.0 - alpha) * ys[last]
            last = i
            cut = x[last] + delta
            i = last + 1
            while i < n:
                if x[i] > cut:
                    break
                if x[i] == x[last]:
                    ys[i] = ys[last]
                    last = i
                i += 1
            i = max(last + 1, i - 1)
            if last >= n - 1:
                break
        res[:] = y - ys
        sc = float(np.sum(np.abs(res))) / n
        if it > nsteps:
            break
        rw[:] = np.abs(res)
        srt = np.sort(rw)
        m1 = n // 2
        if n % 2 == 0:
            m2 = n - m1 - 1
            cmad = 3.0 * (srt[m1] + srt[m2])
        else:
            cmad = 6.0 * srt[m1]
        if cmad < 1e-7 * sc:
            break
        c9, c1 = 0.999 * cmad, 0.001 * cmad
        r = np.abs(res)
        rw[:] = np.where(r <= c1, 1.0, np.where(r <= c9, (1.0 - (r / cmad) ** 2) ** 2, 0.0))
        it += 1
    return ys


def r_lowess(x, y, f=2.0 / 3.0, iter=3, delta=None):
    """stats::lowess: sort by x (stable, like R's order()), delta default 1 % of the x range."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    o = np.argsort(x, kind="stable")
    xs, ysrt = x[o], y[o]
    if delta is None:
        delta = 0.01 * (xs[-1] - xs[0])
    return xs, clowess(xs, ysrt, f, iter, delta)


def r_approx(x, y, xout):
    """stats::approx(x, y, xout) with its defaults: ties averaged, linear, NA outside the range (not reached here)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    ux, inv = np.unique(x, return_inverse=True)
    uy = np.bincount(inv, weights=y) / np.bincount(inv)
    return np.interp(np.asarray(xout, dtype=np.float64), ux, uy)


def gc_normalise(bincount, gc_content, chrom_names):
    """cbs.r:18-25 -> (ratio, lowratio).  chrom_names: bin.chrom column of gc.txt (chr1..chr22, chrX, chrY)."""
    a = np.asarray(bincount, dtype=np.float64) + 1.0
    # cbs.r:13-16: substring(chrom, 4) with X -> 23, Y -> 24; anything else non-numeric is NA under as.numeric and drops
    # out of which(chrom.numeric < 23)
    num = np.array([23 if c == "chrX" else 24 if c == "chrY" else int(c[3:]) if c[3:].isdigit() else 99 for c in chrom_names])
    ratio = a / np.mean(a[num < 23])
    gc = np.asarray(gc_content, dtype=np.float64)
    lx, ly = r_lowess(gc, np.log(ratio), f=0.05)
    z = r_approx(lx, ly, gc)
    return ratio, np.exp(np.log(ratio) - z)
