"""-m gpu: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from helpers import make_case, oracle_tail
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(workdir):
    d = os.path.join(workdir, "gpu_small")
    ref, reads, fa, oix, body = make_case(d, n_pairs=1500, seed=11)
    return dict(dir=d, ref=ref, reads=reads, fa=fa, oix=oix, body=body)


@pytest.fixture(scope="module")
def gpu(case):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True)
    yield api, ix, ctx
    ctx.close()
    ix.close()


def _triples(m):
    return np.stack([m["ref"], m["query"], m["len"]], axis=1).astype(np.uint64)


def test_header(case, gpu):
    api, ix, ctx = gpu
    assert ix.sam_header() == case["oix"].sam_header().encode()


def test_mam_matches_and_sam(case, gpu):
    api, ix, ctx = gpu
    sam, moff, mm = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4, want_matches=True)
    res = ctx.map_batch(case["reads"], want=api.WANT_SAM | api.WANT_MATCHES)
    assert np.array_equal(res.match_off, moff)
    assert np.array_equal(res.matches, _triples(mm))
    assert res.sam == sam                       # byte-exact, same (input) order
    assert ctx.launches > 0


@pytest.mark.parametrize("min_len", [16, 24, 31])
def test_min_len_sweep(case, min_len):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=min_len, nomap=True)
    try:
        sam = case["oix"].map_batch(case["reads"], min_len=min_len, n_threads=4)
        res = ctx.map_batch(case["reads"])
        assert res.sam == sam
    finally:
        ctx.close(); ix.close()


def test_no_nomap_and_n_flag(case):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=False, nucleotides_only=True)
    try:
        sam = case["oix"].map_batch(case["reads"], min_len=20, nomap=False, nucleotides_only=True, n_threads=4)
        assert ctx.map_batch(case["reads"]).sam == sam
    finally:
        ctx.close(); ix.close()


def test_mappability_build(case, gpu):
    api, ix, ctx = gpu
    total = int(case["oix"].sizes[::2].sum())
    body = ctx.build_mappability(total)
    assert np.array_equal(body, case["body"])


def test_tagged_sam_and_tail(case):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, tag_mappability=True)
    try:
        ctx.load_mappability_file(case["fa"] + ".bin/map.bin")
        sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
        exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
        ci = exp["chrominfo"]
        ctx.tail_configure([int(b[2]) for b in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
        res = ctx.map_batch(case["reads"], want=api.WANT_SAM | api.WANT_TAIL)
        assert res.sam == b"".join(exp["tagged"])
        counts, st = ctx.tail_finish()
        chrom, pos = ctx.tail_positions()
        names = case["oix"].descr[::2]
        got = [f"{names[c]} {p}" for c, p in zip(chrom, pos)]
        assert got == exp["positions"]
        assert np.array_equal(counts, exp["counts"])
        assert (st["total_reads"], st["dups_removed"], st["reads_kept"]) == (exp["total"], exp["dups"], exp["kept"])
        assert (st["n_dupe_pairs"], st["n_non_dupe_pairs"]) == (exp["n_dupe"], exp["n_non"])
    finally:
        ctx.close(); ix.close()


def test_double_buffered_submit(case, gpu):
    api, ix, ctx = gpu
    sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
    ctx.submit(0, case["reads"]); ctx.submit(1, case["reads"])
    a = ctx.wait(0); b = ctx.wait(1)
    assert a.sam == sam and b.sam == sam
