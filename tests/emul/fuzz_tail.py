#!/usr/bin/env python
"""tests/emul/fuzz_tail.py -- TEST INFRASTRUCTURE ONLY.

Like fuzz_kernels.py, for the tagged SAM + tail (mappability tags, smashMEM filter, dedupe, varbin) on the emulated
kernels against oracle/tail.py:  python tests/emul/fuzz_tail.py 0 40"""
import os, sys, time, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import shimlib
os.environ["SMASH_B200_LIB"] = shimlib.build()
import numpy as np
from helpers import make_case, oracle_tail
from smash_paper_b200 import api

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
t0 = time.time()
for seed in range(lo, hi):
    rng = np.random.default_rng(777000 + seed)
    kw = dict(n_chrom=int(rng.integers(1, 4)), chrom_len=int(rng.integers(3000, 15000)), n_pairs=int(rng.integers(20, 400)),
              seed=int(rng.integers(1, 10**6)), read_len=int(rng.choice([100, 150, 151, 250])),
              n_families=int(rng.integers(0, 8)), family_len=int(rng.integers(30, 300)), n_long=int(rng.integers(0, 3)),
              n_pad=int(rng.choice([200, 400])), dup_frac=float(rng.choice([0.0, 0.02, 0.2])))
    min_len = int(rng.choice([16, 20, 24]))
    try:
        with tempfile.TemporaryDirectory() as d:
            case_dir = os.path.join(d, "c")
            ref, reads, fa, oix, body = make_case(case_dir, **kw)
            ix = api.Index.open(fa)
            ctx = api.Context(ix, min_len=min_len, nomap=True, tag_mappability=True)
            try:
                ctx.load_mappability_file(fa + ".bin/map.bin")
                sam = oix.map_batch(reads, min_len=min_len, n_threads=2)
                try:
                    exp = oracle_tail(oix, body, sam, case_dir, fa)
                except Exception as e:                       # mappability_tag's "too big" throw: the library must refuse too
                    try:
                        ctx.map_batch(reads, want=api.WANT_SAM | api.WANT_TAIL)
                        print("seed", seed, "oracle threw", repr(e)[:80], "but the library did not"); bad += 1
                    except api.SmashError:
                        pass
                    continue
                ci = exp["chrominfo"]
                ctx.tail_configure([int(b[2]) for b in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
                res = ctx.map_batch(reads, want=api.WANT_SAM | api.WANT_TAIL)
                counts, st = ctx.tail_finish()
                chrom, pos = ctx.tail_positions()
                names = oix.descr[::2]
                got = [f"{names[c]} {p}" for c, p in zip(chrom, pos)]
                ok = (res.sam == b"".join(exp["tagged"]) and got == exp["positions"] and np.array_equal(counts, exp["counts"])
                      and (st["total_reads"], st["dups_removed"], st["reads_kept"]) == (exp["total"], exp["dups"], exp["kept"])
                      and (st["n_dupe_pairs"], st["n_non_dupe_pairs"]) == (exp["n_dupe"], exp["n_non"]))
                if not ok:
                    bad += 1
                    print("MISMATCH seed", seed, kw, min_len, "sam", res.sam == b"".join(exp["tagged"]), "pos", got == exp["positions"],
                          "counts", np.array_equal(counts, exp["counts"]), st, (exp["total"], exp["dups"], exp["kept"], exp["n_dupe"], exp["n_non"]), flush=True)
            finally:
                ctx.close(); ix.close()
    except Exception as e:
        bad += 1
        print("ERROR seed", seed, kw, repr(e)[:300], flush=True)
print("range", lo, hi, "bad", bad, "time", time.time() - t0)
