"""CPU: the oracle (oracle/smash_oracle.c, oracle/tail.py) against the committed golden vectors the
UNMODIFIED reference produced (tests/golden/make_golden.py).  This is what pins the oracle."""
import gzip
import hashlib
import os

import numpy as np
import pytest

from helpers import GOLDEN, golden_lines, golden_variants, load_golden_case
from oracle import oracle as O
from oracle import tail as T

CASES = ["case_basic", "case_adversarial"]
ALL = [(c, v) for c in CASES for v in golden_variants(c)]


@pytest.mark.parametrize("case,variant", ALL, ids=[f"{c}-{v['name']}" for c, v in ALL])
def test_mapout_records_equal_reference(case, variant):
    g = load_golden_case(case)
    hdr, lines = golden_lines(variant["path"])
    sam = g["oix"].map_batch(g["reads"], mode=O.MEM if variant["mode"] == "mem" else O.MAM, min_len=variant["min_len"],
                             nucleotides_only=variant["nuc"], n_threads=4)
    assert g["oix"].sam_header().encode() == hdr
    assert sorted(sam.splitlines(keepends=True)) == lines          # canonical comparison


@pytest.mark.parametrize("case", CASES)
def test_index_files_equal_reference(case, tmp_path):
    """SA/ISA/LCP/text written from the oracle's own builder hash to what the reference's build wrote."""
    g = load_golden_case(case)
    fa = str(tmp_path / "ref.fa")
    open(fa, "wb").write(open(g["fa"], "rb").read())
    g["oix"].save(fa)
    want = dict(l.split() [::-1] for l in open(os.path.join(g["dir"], "index.sha256")))
    for fn, sha in want.items():
        got = hashlib.sha256(open(os.path.join(fa + ".bin", fn), "rb").read()).hexdigest()
        assert got == sha, fn


@pytest.mark.parametrize("case", CASES)
def test_mappability_equals_reference(case):
    g = load_golden_case(case)
    ref = np.frombuffer(gzip.open(os.path.join(g["dir"], "map.bin.gz")).read(), dtype=np.uint8)[2:]   # 2 junk bytes
    assert np.array_equal(g["oix"].mappability(), ref)


def test_tagger_and_varbin_equal_reference():
    g = load_golden_case("case_basic")
    d = g["dir"]
    tagged = gzip.open(os.path.join(d, "tagged.sam.gz")).read().splitlines(keepends=True)
    body = np.frombuffer(gzip.open(os.path.join(d, "map.bin.gz")).read(), dtype=np.uint8)[2:]
    sam = g["oix"].map_batch(g["reads"], min_len=20, n_threads=2)
    hdr = g["oix"].sam_header().encode().splitlines(keepends=True)
    mine = T.tag_lines(hdr + sam.splitlines(keepends=True), g["names"], g["oix"].sizes[::2], body)
    assert sorted(mine) == sorted(tagged)                            # mappability_tag binary output (line order is free)
    rows, nd, nn = T.smash_filter(mine, g["names"])
    pos = T.positions(rows)
    assert "".join(p + "\n" for p in pos).encode() == gzip.open(os.path.join(d, "positions.txt.gz")).read()
    assert open(os.path.join(d, "smash_trailer.txt")).read() == "%d dupes\t%d non-dupes\n" % (nd, nn)
    bins = T.read_table(os.path.join(d, "bins.txt"))
    ci = T.read_chrominfo(os.path.join(d, "chrom_sizes.txt"))
    counts, total, dups, kept = T.varbin(pos, bins, ci)
    assert T.varbin_text(bins, counts, kept).encode() == gzip.open(os.path.join(d, "varbin.txt.gz")).read()   # varbin.py output


def test_bruteforce_spec_agrees_with_faithful_mam():
    """SURVEY App. A.1 (index-free brute force) == the faithful suffix-link MAM on a tiny text."""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    s = acgt[rng.integers(0, 4, size=700)].copy()
    s[300:340] = s[100:140]                                          # a repeat
    oix = O.Index.build(["c1", "c2"], [s[:400], s[400:]])
    for k in range(40):
        p = int(rng.integers(0, 330)); q = bytes(s[p:p + 30]) + bytes(acgt[rng.integers(0, 4, size=12)]) + bytes(s[p + 50:p + 70])
        a = oix.mam(q, min_len=8); b = oix.mam_bruteforce(q, min_len=8)
        assert [tuple(x) for x in a] == [tuple(x) for x in b]
