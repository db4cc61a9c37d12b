#!/usr/bin/env python
"""tests/emul/fuzz_kernels.py -- TEST INFRASTRUCTURE ONLY.

Random small workloads (reference shape, repeat families, read length, min length, mode, flags, seed length) through
the product's kernels EXECUTED ON THE HOST (tests/emul/shimlib.py) against the oracle: SAM bytes and, with WM=1, the
match CSR per read.  Several copies in parallel put the OS scheduler under load, which varies the interleaving of the
emulated lanes.

    python tests/emul/fuzz_kernels.py 0 60            # seeds 0..59
    WM=1 REP=25 python tests/emul/fuzz_kernels.py 61 62   # one seed 25 times, match CSR compared

This is how the duplicate staging of a match in saturated repeat families was found (tests/test_gpu_parity.py::
test_saturated_repeat_family_reports_each_match_once)."""
import os, sys, time, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import shimlib
os.environ["SMASH_B200_LIB"] = shimlib.build()
import numpy as np
from helpers import make_case
from oracle import oracle as O
from smash_paper_b200 import api, synth
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
t00=time.time()
REP=int(os.environ.get('REP','1'))
for seed in [x for x in range(lo, hi) for _ in range(REP)]:
    rng = np.random.default_rng(9000+seed)
    kw = dict(n_chrom=int(rng.integers(1,4)), chrom_len=int(rng.integers(1500, 12000)), n_pairs=int(rng.integers(5, 60)),
              seed=int(rng.integers(1, 10**6)), read_len=int(rng.choice([24, 36, 50, 100, 150, 151, 250, 400])),
              n_families=int(rng.integers(0, 8)), family_len=int(rng.integers(30, 300)), n_long=int(rng.integers(0, 3)),
              n_highcopy=int(rng.integers(0, 2)), highcopy_copies=int(rng.integers(20, 120)), n_pad=int(rng.choice([0, 50, 200])),
              sub_rate=float(rng.choice([0.0, 0.0075, 0.03])), z_rate=float(rng.choice([0.0, 0.01, 0.05])))
    if os.environ.get('READ_LEN'): kw['read_len'] = int(os.environ['READ_LEN']); kw['n_pairs'] = min(kw['n_pairs'], 12)   # e.g. 1500: the long-read kernel
    min_len = int(rng.integers(8, 32)); mode = int(rng.choice([api.MODE_MAM, api.MODE_MAM, api.MODE_MUM, api.MODE_MEM]))
    if os.environ.get('MODE'): mode = {'mam': api.MODE_MAM, 'mum': api.MODE_MUM, 'mem': api.MODE_MEM}[os.environ['MODE']]
    nomap = bool(rng.integers(0, 2)); nuc = bool(rng.integers(0, 2)); seed_k = int(rng.choice([0, 0, 5, 7, 9]))
    try:
        with tempfile.TemporaryDirectory() as d:
            ref, reads, fa, oix, body = make_case(os.path.join(d, "c"), **kw)
            ix = api.Index.open(fa)
            ctx = api.Context(ix, device=0, mode=mode, min_len=min_len, nomap=nomap, nucleotides_only=nuc, seed_k=seed_k)
            try:
                res = ctx.map_batch(reads, want=api.WANT_SAM | (api.WANT_MATCHES if os.environ.get('WM')=='1' else 0))
                omode = {api.MODE_MAM: O.MAM, api.MODE_MUM: O.MUM, api.MODE_MEM: O.MEM}[mode]
                exp = oix.map_batch(reads, mode=omode, min_len=min_len, nomap=nomap, nucleotides_only=nuc)
                ok = res.sam == exp
            except api.SmashError as e:
                ok = "keeps at most" in str(e)      # documented limit (> 64 MAMs per read)
                if ok: print("seed", seed, "limit:", str(e)[:80])
            finally:
                ctx.close(); ix.close()
        if not ok:
            a=res.sam.splitlines(); b=exp.splitlines(); sa=set(a); sb=set(b)
            oa=[l for l in a if l not in sb]; ob=[l for l in b if l not in sa]
            print("only gpu", len(oa), "only oracle", len(ob), "lines", len(a), len(b))
            if res.matches is not None:
                for i in range(reads.n):
                    gm=[tuple(int(x) for x in m) for m in res.matches[res.match_off[i]:res.match_off[i+1]]]
                    q=bytes(reads.seq[reads.seq_off[i]:reads.seq_off[i+1]]).lower()
                    if nuc: q=bytes(c if c in b"acgt" else ord("~") for c in q)
                    om=[tuple(int(x) for x in m) for m in oix.mam(q, min_len)]
                    if gm!=om: print("read", i, "gpu", gm, "oracle", om)
            for l in oa[:3]: print("G", l.split(b"\t")[:9], l.split(b"\t")[11:])
            for l in ob[:3]: print("O", l.split(b"\t")[:9], l.split(b"\t")[11:])
            bad += 1
            print("MISMATCH seed", seed, kw, "min_len", min_len, "mode", mode, "nomap", nomap, "nuc", nuc, "seed_k", seed_k, flush=True)
    except Exception as e:
        bad += 1
        print("ERROR seed", seed, kw, min_len, mode, repr(e)[:300], flush=True)
print("range", lo, hi, "bad", bad, "time", time.time()-t00)
