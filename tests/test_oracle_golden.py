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


@pytest.mark.parametrize("case", ["case_basic", "case_tail"])
def test_smash_filter_equals_unmodified_smashmem_script(case):
    """oracle/tail.py:smash_filter against what the UNMODIFIED /root/reference/smashMEM.py printed for the same tagged SAM
    (tests/golden/make_golden_smash.py: the script run over a test-only pysam stand-in and the restated samtools name
    sort).  case_tail: read-2 hits at 500/9999/10000/10001/25000 bp from a read-1 hit, cross-pair duplicates far apart,
    read names whose name order is not the input order, chrM/_gl000 hits, excess-mappability failures."""
    d = os.path.join(GOLDEN, case)
    tagged = gzip.open(os.path.join(d, "tagged.sam.gz")).read().splitlines(keepends=True)
    names = [ln.split(b"\t")[1][3:].decode() for ln in tagged if ln.startswith(b"@SQ")]
    rows, nd, nn = T.smash_filter(tagged, names)
    assert T.smash_text(rows, nd, nn).encode() == gzip.open(os.path.join(d, "smash.txt.gz")).read()
    pos = T.positions(rows)
    assert "".join(p + "\n" for p in pos).encode() == gzip.open(os.path.join(d, "positions.txt.gz")).read()
    bins = T.read_table(os.path.join(d, "bins.txt"))
    ci = T.read_chrominfo(os.path.join(d, "chrom_sizes.txt"))
    counts, total, dups, kept = T.varbin(pos, bins, ci)
    assert T.varbin_text(bins, counts, kept).encode() == gzip.open(os.path.join(d, "varbin.txt.gz")).read()   # varbin.py output
    if case == "case_tail":
        assert nd > 10 and dups > 0 and total > kept                       # the stress cases are really in there
        in_order = [ln.split(b"\t")[0] for ln in gzip.open(os.path.join(d, "reads.sam.gz")).read().splitlines()][::2]
        assert in_order != [ln.split(b"\t")[0] for ln in T.name_sort_lines([n + b"\t77\t" for n in in_order])]


def test_samtools_name_order():
    """strnum_cmp (samtools 0.1.x bam_sort.c): digit runs as numbers, fewer leading zeros sorts later, ties by mate."""
    c = T.strnum_cmp
    assert c(b"r9", b"r10") < 0 and c(b"r10", b"r9") > 0 and c(b"r10", b"r10") == 0
    assert c(b"r010", b"r10") < 0 and c(b"r10", b"r010") > 0
    assert c(b"x:12:5", b"x:2:50") > 0 and c(b"a1b", b"a1c") < 0 and c(b"a", b"a1") < 0 and c(b"a2", b"ab") < 0
    lines = [b"b\t141\tx\n", b"a10\t77\tx\n", b"b\t77\tx\n", b"a9\t141\tx\n", b"a9\t77\tx\n"]
    assert T.name_sort_lines(lines) == [lines[4], lines[3], lines[1], lines[2], lines[0]]


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
