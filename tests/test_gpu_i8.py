"""-m gpu: parity at the 8-BYTE index width -- the width bench.py's config 2 (hg19-shaped, N = 6.19e9) runs at.

The reference ships three integer widths (size.h:9-22, Makefile:16-23) and switches binaries by FASTA size
(mummer.cpp:156-183); `mummer-long` run directly on a small FASTA writes and reads `rc1.i8.index.*` files with the
very same code it uses at hg19 scale.  Everything here is checked against that UNMODIFIED binary run live
(oracle/_ref/mummer-long), on a reference small enough for its qsufsort build: index files byte for byte, map.bin,
SAM records (MAM at several seed lengths -- 8-byte seed table, 4+4 pre-filter and split search included -- MUM and
MEM), mappability_tag's L/R tags, the positions list and the bin counts.
"""
import filecmp
import glob
import os
import shutil

import numpy as np
import pytest

from helpers import oracle_tail
from oracle import oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built")]

I8_FILES = ["rc1.ref.bin", "rc1.ref.seq.bin", "rc1.i8.index.bin", "rc1.i8.index.sa.bin", "rc1.i8.index.isa.bin",
            "rc1.i8.index.lcp.vec.bin", "rc1.i8.index.lcp.m.bin"]


@pytest.fixture(scope="module")
def i8(workdir):
    from smash_paper_b200 import synth
    d = os.path.join(workdir, "i8")
    os.makedirs(d, exist_ok=True)
    ref = synth.make_reference([("chr1", 1_300_000), ("chr2", 900_000), ("chrX", 400_000)], seed=41, n_families=40, n_long=4)
    fa = os.path.join(d, "ref.fa")
    synth.write_fasta(ref, fa)
    synth.write_index_side_files(ref, fa)
    synth.write_fixed_bins(ref, os.path.join(d, "bins.txt"), width=20000)
    reads = synth.make_reads(ref, 4000, seed=42)
    synth.write_sam(reads, os.path.join(d, "reads.sam"))
    O.ref_build_index(fa, long_ints=True, mappability=True)          # mummer-long: rc1.i8.* + map.bin
    assert os.path.exists(fa + ".bin/rc1.i8.index.sa.bin") and not os.path.exists(fa + ".bin/rc1.i4.index.sa.bin")
    oix = O.Index.load(fa)
    assert oix.w == 8 and oix.sa.dtype.itemsize == 8
    body = np.fromfile(fa + ".bin/map.bin", dtype=np.uint8)[2:]
    return dict(dir=d, ref=ref, fa=fa, reads=reads, oix=oix, body=body)


def _ref_lines(i8, extra=()):
    return O.ref_map(i8["fa"], os.path.join(i8["dir"], "reads.sam"), i8["dir"], threads=4, extra=extra, long_ints=True)


def test_i8_index_files_written_by_the_gpu_builder_are_byte_identical(i8, workdir):
    """smash_ctx_create_from_text(w=8) + smash_ctx_save_index == what `mummer-long -rcref <fa> dummy` and
    `-mappability` left behind (longSA.cpp:179-190, fasta.cpp:215-236)."""
    from smash_paper_b200 import api
    oix = i8["oix"]
    d = os.path.join(workdir, "i8_saved")
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "ref.fa")
    shutil.copy(i8["fa"], fa)
    ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=8, chunk_cap=700_000)
    try:
        sa, isa, vec, m = ctx.copy_index(oix.N, 8)
        assert sa.dtype == np.uint64 and np.array_equal(sa, oix.sa) and np.array_equal(isa, oix.isa)
        assert np.array_equal(vec, oix.lcp_vec)
        body = ctx.build_mappability(int(oix.sizes[::2].sum()))
        assert np.array_equal(body, i8["body"])
        ctx.save_index(fa, with_mappability=True)
    finally:
        ctx.close()
    assert sorted(os.path.basename(p) for p in glob.glob(fa + ".bin/rc1.*")) == sorted(I8_FILES)
    for f in I8_FILES:
        assert filecmp.cmp(os.path.join(fa + ".bin", f), os.path.join(i8["fa"] + ".bin", f), shallow=False), f
    assert np.array_equal(np.fromfile(fa + ".bin/map.bin", dtype=np.uint8)[2:], i8["body"])
    # and the saved files open again as an 8-byte index that maps the same
    from smash_paper_b200 import api as A
    ix = A.Index.open(fa)
    try:
        assert ix.int_width == 8
        c2 = A.Context(ix, min_len=20, nomap=True)
        try:
            assert c2.map_batch(i8["reads"]).sam == oix.map_batch(i8["reads"], min_len=20, n_threads=4)
        finally:
            c2.close()
    finally:
        ix.close()


@pytest.mark.parametrize("seed_k", [0, 9, 11, 12])
def test_i8_mam_records_equal_mummer_long(i8, seed_k):
    """smash_index_open on mummer-long's own files, 8-byte SA + 8-byte seed entries; seed_k 9..12 turn on the
    4+4 pre-filter and the split search (k_mam_search parks, k_mam_verify extends) exactly as at hg19 scale."""
    from smash_paper_b200 import api
    hdr, lines = _ref_lines(i8)
    ix = api.Index.open(i8["fa"])
    assert ix.int_width == 8
    ctx = api.Context(ix, min_len=20, nomap=True, seed_k=seed_k)
    try:
        assert ix.sam_header() == hdr
        res = ctx.map_batch(i8["reads"], want=api.WANT_SAM | api.WANT_MATCHES)
        assert sorted(res.sam.splitlines(keepends=True)) == lines                     # what mummer-long printed
        sam, moff, mm = i8["oix"].map_batch(i8["reads"], min_len=20, n_threads=4, want_matches=True)
        assert res.sam == sam                                                          # input order, via the oracle
        assert np.array_equal(res.match_off, moff)
        assert np.array_equal(res.matches, np.stack([mm["ref"], mm["query"], mm["len"]], axis=1).astype(np.uint64))
    finally:
        ctx.close(); ix.close()


@pytest.mark.parametrize("flags,mode,min_len", [(["-mum"], "mum", 20), (["-maxmatch"], "mem", 20), (["-l", "16"], "mam", 16),
                                                (["-maxmatch", "-l", "17"], "mem", 17)])
def test_i8_other_modes_equal_mummer_long(i8, flags, mode, min_len):
    from smash_paper_b200 import api
    hdr, lines = _ref_lines(i8, extra=flags)
    ix = api.Index.open(i8["fa"])
    ctx = api.Context(ix, min_len=min_len, nomap=True, mode={"mum": api.MODE_MUM, "mem": api.MODE_MEM, "mam": api.MODE_MAM}[mode])
    try:
        res = ctx.map_batch(i8["reads"])
        assert sorted(res.sam.splitlines(keepends=True)) == lines
    finally:
        ctx.close(); ix.close()


@pytest.mark.parametrize("seed_k", [0, 11])
def test_i8_tagged_sam_positions_and_bin_counts(i8, seed_k):
    """L/R tags against the unmodified mappability_tag binary run on mummer-long's records; positions list and bin
    counts against the (pinned) tail oracle fed with those reference-made lines."""
    from smash_paper_b200 import api
    hdr, lines = _ref_lines(i8)
    allsam = os.path.join(i8["dir"], "all_i8.sam")
    in_order = sorted(lines, key=lambda l: (l.split(b"\t")[0], int(l.split(b"\t")[1]) & 128, _hi(l)))
    open(allsam, "wb").write(hdr + b"".join(in_order))
    tagged_ref = O.ref_mappability_tag(i8["fa"], allsam)
    tagged_ref_lines = [l for l in tagged_ref.splitlines(keepends=True) if not l.startswith(b"@")]
    ix = api.Index.open(i8["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, tag_mappability=True, seed_k=seed_k)
    try:
        ctx.load_mappability_file(i8["fa"] + ".bin/map.bin")
        exp = oracle_tail(i8["oix"], i8["body"], b"".join(in_order), i8["dir"], i8["fa"])
        assert exp["tagged"] == tagged_ref_lines                   # the tail oracle's tagger == the reference's
        ci = exp["chrominfo"]
        ctx.tail_configure([int(b[2]) for b in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
        res = ctx.map_batch(i8["reads"], want=api.WANT_SAM | api.WANT_TAIL)
        assert res.sam == b"".join(tagged_ref_lines)               # byte-exact, record order = read order, HI order
        counts, st = ctx.tail_finish()
        chrom, pos = ctx.tail_positions()
        names = i8["oix"].descr[::2]
        assert [f"{names[c]} {p}" for c, p in zip(chrom, pos)] == exp["positions"]
        assert np.array_equal(counts, exp["counts"])
        assert (st["total_reads"], st["dups_removed"], st["reads_kept"]) == (exp["total"], exp["dups"], exp["kept"])
        assert (st["n_dupe_pairs"], st["n_non_dupe_pairs"]) == (exp["n_dupe"], exp["n_non"])
    finally:
        ctx.close(); ix.close()


def _hi(line):
    for f in line.rstrip(b"\n").split(b"\t")[11:]:
        if f.startswith(b"HI:i:"):
            return int(f[5:])
    return 0
