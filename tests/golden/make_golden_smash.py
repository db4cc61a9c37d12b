#!/usr/bin/env python
"""Golden vectors of the smashMEM.py stage, produced by the UNMODIFIED reference script.

    python tests/golden/make_golden_smash.py          # dev container only (needs /root/reference and oracle/_ref)

`/root/reference/smashMEM.py` parses under python3 and needs two things that are not installable here: the `pysam`
module and a name-sorted BAM made by `samtools sort -n` (smash_mapping.sh:23-26).  It is run here as a subprocess,
byte for byte as shipped, with
  * tests/stubs/pysam.py on PYTHONPATH (a test-only stand-in for the accessors the script touches, semantics taken from
    the pysam documentation), reading SAM TEXT instead of BAM, and
  * its input name-sorted by oracle/tail.py:name_sort_lines (samtools 0.1.x strnum_cmp order, restated).
What is committed per case:
  smash.txt.gz        the script's stdout for argv `<file> 0 0 10000 4` (header row, one row per kept hit, trailer)
  positions.txt.gz    `awk '{print $4, $5}' | perl -ne 'print if /^chr(\\d+|[XY]) \\d+$/'` of it (smash_mapping.sh:29)
and for case_tail (a workload built to stress this stage: read-2 hits inside and outside the 10 kb window of a read-1
hit, cross-pair duplicates far apart in the input, read names whose name order is NOT the input order, chrM and
*_gl000* chromosomes, hits that fail the excess-mappability filter) the whole chain from the reference binaries:
  ref.fa.gz reads.sam.gz bins.txt chrom_sizes.txt map.bin.gz tagged.sam.gz varbin.txt.gz stats.txt
"""
import gzip
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import tail as T  # noqa: E402
from smash_paper_b200 import synth  # noqa: E402

REF = "/root/reference"
STUBS = os.path.join(ROOT, "tests", "stubs")
_POSRE = re.compile(rb"^chr(\d+|[XY]) \d+$")


def gz_write(path, data: bytes):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(data)


def run_smashmem(tagged: bytes, workdir):
    """tagged SAM (header + records, any order) -> stdout of the unmodified smashMEM.py."""
    lines = tagged.splitlines(keepends=True)
    hdr = [ln for ln in lines if ln.startswith(b"@")]
    rec = T.name_sort_lines([ln for ln in lines if not ln.startswith(b"@")])
    path = os.path.join(workdir, "namesort.sam")
    open(path, "wb").write(b"".join(hdr + rec))
    env = dict(os.environ, PYTHONPATH=STUBS)
    r = subprocess.run([sys.executable, os.path.join(REF, "smashMEM.py"), path, "0", "0", "10000", "4"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    return r.stdout


def positions_of(smash_txt: bytes):
    out = []
    for ln in smash_txt.splitlines():
        f = ln.split()                                        # awk's default field splitting
        if len(f) >= 5:
            s = f[3] + b" " + f[4]
            if _POSRE.match(s):
                out.append(s + b"\n")
    return b"".join(out)


def tail_reference(seed=77):
    sizes = [("chr1", 60000), ("chr2", 45000), ("chrX", 30000), ("chrM", 8000), ("chrUn_gl000220", 9000)]
    return synth.make_reference(sizes, seed=seed, n_pad=300, n_families=8, family_len=120, family_copies=3, n_long=2, long_len=400)


def tail_reads(ref, seed=78):
    rng = np.random.default_rng(seed)
    base = synth.make_reads(ref, 360, read_len=150, seed=seed, dup_frac=0.04)
    seqs = [bytes(base.seq[base.seq_off[i]:base.seq_off[i + 1]]) for i in range(base.n)]
    g = [bytes(s) for s in ref.seqs]
    acgt = b"ACGT"

    def rnd(n):
        return bytes(acgt[int(x)] for x in rng.integers(0, 4, size=n))

    def rc(s):
        return synth._COMP[np.frombuffer(s, dtype=np.uint8)[::-1]].tobytes()

    pairs = [(seqs[2 * i], seqs[2 * i + 1]) for i in range(base.n // 2)]
    c1, c2, cx, cm, cu = g
    # read-2 hit near / far from a read-1 hit on the same chromosome (smashMEM.py:193-208): 9999 is near, 10000 is not
    for d in (500, 9999, 10000, 10001, 25000):
        a = 5000
        pairs.append((c1[a:a + 60] + rnd(30) + c2[7000:7060], c1[a + d:a + d + 70] + rnd(20) + cx[4000:4060]))
        pairs.append((rc(c1[a + 100:a + 170]) + rnd(80), rnd(40) + rc(c1[a + 100 + d:a + 100 + d + 80]) + rnd(30)))
    # the same key under different names, far apart in the input (first-wins dedupe, smashMEM.py:217-228) ...
    dup_a = (c2[12000:12070] + rnd(15) + cx[9000:9065], c1[33000:33075] + rnd(75))
    # ... and pairs that share only part of the key (not duplicates)
    dup_b = (dup_a[0], c1[33000:33075] + rnd(10) + c2[30000:30065])
    pairs.insert(3, dup_a)
    pairs.insert(200, dup_b)
    pairs.append(dup_a)
    pairs.append((dup_a[0], dup_a[1][:75] + rnd(75)))         # same hits again, different junk
    # read 1 without a passing hit, read 2 with hits; chrM / *_gl000* hits (dropped by the perl filter / varbin)
    pairs.append((rnd(150), c2[20000:20080] + rnd(70)))
    pairs.append((cm[1000:1090] + rnd(60), cu[2000:2085] + rnd(65)))
    pairs.append((c1[40000:40070] + rnd(10) + cm[3000:3070], rnd(150)))
    n_free = len(pairs)
    # equal position strings on consecutive kept lines (varbin.py:56-58 drops every line whose position STRING equals the
    # previous kept line's, whatever the chromosome): these pairs get names that are adjacent in name order; a chrM hit
    # between two of them is removed by the perl filter first, so they still meet
    fixed = [(b"HWI:0:1:1:1", (c1[15000:15080] + rnd(70), rnd(150))),
             (b"HWI:0:1:1:2", (c2[15000:15080] + rnd(70), rnd(150))),
             (b"HWI:0:1:1:3", (cm[2000:2080] + rnd(70), rnd(150))),
             (b"HWI:0:1:1:4", (rnd(70) + cx[15000:15080], rnd(150))),
             (b"HWI:0:1:1:5", (c1[15001:15081] + rnd(70), rnd(150))),
             (b"HWI:0:1:01:6", (rnd(150), c2[15001:15081] + rnd(70)))]
    pairs += [p for _, p in fixed]
    n_pairs = len(pairs)
    # names whose samtools name order differs from the input order: lane:tile:x:y style fields of varying width,
    # leading zeros, and a few names that differ only in zero padding
    names = []
    used = set(nm for nm, _ in fixed)
    while len(names) < n_free:
        k = len(names)
        if k % 37 == 5:
            nm = b"HWI:%d:%04d:%d:%d" % (1 + k % 2, 1101 + (k * 7) % 13, int(rng.integers(1, 20000)), int(rng.integers(1, 99999)))
        elif k % 37 == 6:
            nm = b"HWI:%d:%d:%05d:%d" % (1 + k % 2, 1101 + (k * 7) % 13, int(rng.integers(1, 20000)), int(rng.integers(1, 99999)))
        else:
            nm = b"HWI:%d:%d:%d:%d" % (1 + k % 2, 1101 + (k * 7) % 13, int(rng.integers(1, 20000)), int(rng.integers(1, 99999)))
        if nm not in used:
            used.add(nm); names.append(nm)
    names += [nm for nm, _ in fixed]
    order = rng.permutation(n_pairs)
    pairs = [pairs[i] for i in order]
    names = [names[i] for i in order]
    nl, sl, ql, fl = [], [], [], []
    for nm, (s1, s2) in zip(names, pairs):
        for s, f in ((s1, 77), (s2, 141)):
            nl.append(nm); sl.append(s); fl.append(f)
            ql.append(bytes(33 + int(x) for x in rng.integers(2, 40, size=len(s))))

    def blob(lst):
        off = np.zeros(len(lst) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in lst])
        return np.frombuffer(b"".join(lst), dtype=np.uint8).copy(), off

    nb, no = blob(nl); sb, so = blob(sl); qb, _ = blob(ql)
    return synth.ReadBatch(names=nb, name_off=no, seq=sb, qual=qb, seq_off=so, flags=np.array(fl, dtype=np.uint16),
                           opt=np.zeros(0, dtype=np.uint8), opt_off=np.zeros(len(nl) + 1, dtype=np.int64))


def case_tail():
    out = os.path.join(HERE, "case_tail")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    ref = tail_reference()
    reads = tail_reads(ref)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "ref.fa")
        synth.write_fasta(ref, fa)
        synth.write_index_side_files(ref, fa)
        synth.write_sam(reads, os.path.join(d, "reads.sam"))
        O.ref_build_index(fa, mappability=True)
        hdr, lines = O.ref_map(fa, os.path.join(d, "reads.sam"), d)
        # smash_mapping.sh:23: header once, records with the perl name rewrite, through mappability_tag
        lines = [re.sub(rb"^(\S+?)/\S+/\d+", rb"\1", ln) for ln in lines]
        allsam = os.path.join(d, "all.sam")
        open(allsam, "wb").write(hdr + b"".join(lines))
        tagged = O.ref_mappability_tag(fa, allsam)
        smash = run_smashmem(tagged, d)
        pos = positions_of(smash)
        open(os.path.join(d, "positions.txt"), "wb").write(pos)
        synth.write_fixed_bins(ref, os.path.join(d, "bins.txt"), width=5000)
        subprocess.run([sys.executable, os.path.join(REF, "varbin.py"), os.path.join(d, "positions.txt"), os.path.join(d, "bins.txt"),
                        os.path.join(d, "varbin.txt"), os.path.join(d, "stats.txt"), fa + ".bin/chrom_sizes.txt"],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)          # python3: dies at varbin.py:113 AFTER varbin.txt is complete
        gz_write(os.path.join(out, "ref.fa.gz"), open(fa, "rb").read())
        gz_write(os.path.join(out, "reads.sam.gz"), open(os.path.join(d, "reads.sam"), "rb").read())
        gz_write(os.path.join(out, "map.bin.gz"), open(fa + ".bin/map.bin", "rb").read())
        gz_write(os.path.join(out, "tagged.sam.gz"), tagged)
        gz_write(os.path.join(out, "smash.txt.gz"), smash)
        gz_write(os.path.join(out, "positions.txt.gz"), pos)
        gz_write(os.path.join(out, "varbin.txt.gz"), open(os.path.join(d, "varbin.txt"), "rb").read())
        shutil.copy(os.path.join(d, "bins.txt"), os.path.join(out, "bins.txt"))
        shutil.copy(fa + ".bin/chrom_sizes.txt", os.path.join(out, "chrom_sizes.txt"))
    rows = smash.count(b"\n") - 2
    print("case_tail:", reads.n // 2, "pairs,", rows, "smash rows,", pos.count(b"\n"), "positions;", smash.splitlines()[-1].decode())


def case_basic():
    d = os.path.join(HERE, "case_basic")
    tagged = gzip.open(os.path.join(d, "tagged.sam.gz")).read()
    with tempfile.TemporaryDirectory() as tmp:
        smash = run_smashmem(tagged, tmp)
    gz_write(os.path.join(d, "smash.txt.gz"), smash)
    pos = positions_of(smash)
    old = gzip.open(os.path.join(d, "positions.txt.gz")).read()
    print("case_basic:", smash.count(b"\n") - 2, "smash rows; positions.txt", "unchanged" if old == pos else "CHANGED", ";", smash.splitlines()[-1].decode())
    gz_write(os.path.join(d, "positions.txt.gz"), pos)
    open(os.path.join(d, "smash_trailer.txt"), "wb").write(smash.splitlines()[-1] + b"\n")


if __name__ == "__main__":
    assert O.have_reference() and os.path.exists(os.path.join(REF, "smashMEM.py"))
    case_basic()
    case_tail()
