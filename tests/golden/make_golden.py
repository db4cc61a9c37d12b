#!/usr/bin/env python
"""Generate the committed golden fixtures with the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the dev container only:

    python tests/golden/make_golden.py

Every fixture directory holds the inputs (ref.fa.gz, reads.sam.gz, bins/chrom_sizes) and what the
reference itself printed for them:
  mapout_<variant>.sam.gz   header + byte-sorted record lines of `mummer -rcref -qthreads 2 -nomap
                            -samin -samout [flags]` (smash_mapping.sh:19)
  index.sha256              sha256 of every <fa>.bin/rc1.* file the reference's build wrote
  map.bin.gz                `mummer -rcref -mappability` output (index_setup.sh:22)
  tagged.sam.gz             `mappability_tag` output (smash_mapping.sh:23)
  varbin.txt.gz             `varbin.py` output (binning.sh:36) for positions.txt
  positions.txt.gz          smashMEM.py stage -- NOT produced by the reference (pysam/samtools are
                            absent): written by oracle/tail.py from tagged.sam ("parity unpinned")
"""
import gzip
import hashlib
import os
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


def gz_write(path, data: bytes):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(data)


def sha_dir(d):
    out = []
    for fn in sorted(os.listdir(d)):
        if fn.startswith("rc1."):
            data = open(os.path.join(d, fn), "rb").read()
            if fn.endswith(".i4.index.lcp.m.bin"):
                # item_t {size_t idx; uint32 val; <4 pad bytes>}: the pad is uninitialised heap memory in
                # the reference (longSA.cpp:43-56) -> hash with the pad zeroed, which is what we write
                a = np.frombuffer(data, dtype=np.uint8).reshape(-1, 16).copy()
                a[:, 12:] = 0
                data = a.tobytes()
            out.append(f"{hashlib.sha256(data).hexdigest()}  {fn}\n")
    return "".join(out)


def adversarial_reference(seed=21):
    """High-copy repeats (bucket overflow / expand_link threshold), tandem repeats, a palindrome,
    >255 bp repeats, IUPAC letters, and NO N padding (so matches can touch chromosome starts)."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = synth._COMP
    chroms = []
    for ci, L in enumerate([16000, 15000, 9000]):
        s = acgt[rng.integers(0, 4, size=L)].copy()
        chroms.append(s)
    unit = acgt[rng.integers(0, 4, size=48)]
    for _ in range(260):                                   # high-copy family
        c = int(rng.integers(0, 3)); p = int(rng.integers(100, len(chroms[c]) - 200))
        chroms[c][p:p + 48] = unit if rng.random() < 0.7 else comp[unit[::-1]]
    tr = np.frombuffer(b"ACGTTGCA" * 120, dtype=np.uint8)   # tandem repeat (960 bp, LCP >= 255)
    chroms[0][3000:3000 + len(tr)] = tr
    half = acgt[rng.integers(0, 4, size=150)]
    pal = np.concatenate([half, comp[half[::-1]]])          # reverse-complement palindrome
    chroms[1][5000:5300] = pal
    big = acgt[rng.integers(0, 4, size=420)]                # >255 bp exact repeat, two strands
    chroms[0][9000:9420] = big; chroms[2][2000:2420] = comp[big[::-1]]
    chroms[2][6000:6010] = np.frombuffer(b"RYKMNNNNSW", dtype=np.uint8)   # IUPAC + N inside
    return synth.Reference(["chrA", "chrB", "chr_gl000_x"], chroms)


def adversarial_reads(ref, seed=22):
    rng = np.random.default_rng(seed)
    base = synth.make_reads(ref, 220, read_len=150, seed=seed, sub_rate=0.004, z_rate=0.004, dup_frac=0.02)
    seqs = [bytes(base.seq[base.seq_off[i]:base.seq_off[i + 1]]) for i in range(base.n)]
    quals = [bytes(base.qual[base.seq_off[i]:base.seq_off[i + 1]]) for i in range(base.n)]
    names = [bytes(base.names[base.name_off[i]:base.name_off[i + 1]]) for i in range(base.n)]
    flags = list(base.flags)
    opts = [b""] * base.n
    g = [bytes(s) for s in ref.seqs]
    acgt = b"ACGT"

    def rnd(n):
        return bytes(acgt[int(x)] for x in rng.integers(0, 4, size=n))

    def add_pair(s1, s2, name, f1=77, f2=141, o1=b"", o2=b""):
        for s, f, o in ((s1, f1, o1), (s2, f2, o2)):
            seqs.append(s); quals.append(bytes(33 + int(x) for x in rng.integers(2, 40, size=len(s))))
            names.append(name); flags.append(f); opts.append(o)

    # matches that start before a chromosome start (pos < 0 erasure, query.cpp:239-246)
    add_pair(rnd(60) + g[1][:45] + rnd(45), rnd(30) + g[2][:70] + rnd(50), b"edge_start")
    add_pair(g[0][-50:] + rnd(100), synth._COMP[np.frombuffer(g[1][:60], dtype=np.uint8)[::-1]].tobytes() + rnd(90), b"edge_end")
    # reads shorter than min_len, lower case, optional fields, a whole-read match, N/IUPAC bases
    add_pair(rnd(12), g[0][500:519], b"short")
    add_pair(g[0][1000:1150].lower(), g[1][2000:2150], b"lower", o1=b"\tXX:Z:opt1\tYY:i:7", o2=b"\tZZ:Z:hello")
    add_pair(g[2][5990:6030] + rnd(110), g[0][3000:3150], b"iupac_tandem")
    add_pair(g[1][5000:5150], g[1][5150:5300], b"palindrome")
    add_pair(g[0][9000:9150], g[0][9100:9250], b"bigrepeat")
    # name quirks: a name that itself ends in :0 / :1 with flags that carry no mate bits
    add_pair(g[0][200:350], g[1][300:450], b"plain:1", f1=0, f2=0)
    add_pair(g[0][400:550], g[1][600:750], b"quirk:0", f1=4, f2=4)
    # odd read count: a final unpaired read
    seqs.append(g[1][7000:7150]); quals.append(b"I" * 150); names.append(b"last_unpaired"); flags.append(77); opts.append(b"")

    def blob(lst):
        off = np.zeros(len(lst) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in lst])
        return np.frombuffer(b"".join(lst), dtype=np.uint8).copy(), off

    nb, no = blob(names); sb, so = blob(seqs); qb, _ = blob(quals); ob, oo = blob(opts)
    return synth.ReadBatch(names=nb, name_off=no, seq=sb, qual=qb, seq_off=so,
                           flags=np.array(flags, dtype=np.uint16), opt=ob, opt_off=oo)


def run_case(name, ref, reads, variants, tail):
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "ref.fa")
        synth.write_fasta(ref, fa)
        synth.write_index_side_files(ref, fa)
        synth.write_sam(reads, os.path.join(d, "reads.sam"))
        O.ref_build_index(fa, mappability=True)
        gz_write(os.path.join(out, "ref.fa.gz"), open(fa, "rb").read())
        gz_write(os.path.join(out, "reads.sam.gz"), open(os.path.join(d, "reads.sam"), "rb").read())
        open(os.path.join(out, "index.sha256"), "w").write(sha_dir(fa + ".bin"))
        gz_write(os.path.join(out, "map.bin.gz"), open(fa + ".bin/map.bin", "rb").read())
        for vname, flags in variants.items():
            hdr, lines = O.ref_map(fa, os.path.join(d, "reads.sam"), d, extra=flags)
            gz_write(os.path.join(out, f"mapout_{vname}.sam.gz"), hdr + b"".join(lines))
            print(name, vname, len(lines), "records")
        if tail:
            hdr, lines = O.ref_map(fa, os.path.join(d, "reads.sam"), d)
            allsam = os.path.join(d, "all.sam")
            # name order == input order for r%09d names: feed the tagger in read order
            open(allsam, "wb").write(hdr + b"".join(sorted(lines, key=lambda l: (l.split(b"\t")[0], int(l.split(b"\t")[1]) & 128))))
            tagged = O.ref_mappability_tag(fa, allsam)
            gz_write(os.path.join(out, "tagged.sam.gz"), tagged)
            rows, nd, nn = T.smash_filter(tagged.splitlines(keepends=True), ref.names)
            pos = T.positions(rows)
            ptxt = "".join(p + "\n" for p in pos)
            open(os.path.join(d, "positions.txt"), "w").write(ptxt)
            gz_write(os.path.join(out, "positions.txt.gz"), ptxt.encode())
            synth.write_fixed_bins(ref, os.path.join(d, "bins.txt"), width=5000)
            shutil.copy(os.path.join(d, "bins.txt"), os.path.join(out, "bins.txt"))
            shutil.copy(fa + ".bin/chrom_sizes.txt", os.path.join(out, "chrom_sizes.txt"))
            subprocess.run([sys.executable, "/root/reference/varbin.py", os.path.join(d, "positions.txt"),
                            os.path.join(d, "bins.txt"), os.path.join(d, "varbin.txt"), os.path.join(d, "stats.txt"),
                            fa + ".bin/chrom_sizes.txt"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            gz_write(os.path.join(out, "varbin.txt.gz"), open(os.path.join(d, "varbin.txt"), "rb").read())
            open(os.path.join(out, "smash_trailer.txt"), "w").write("%d dupes\t%d non-dupes\n" % (nd, nn))


def main():
    assert O.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    sizes = [("chr1", 20000), ("chr2", 20137), ("chr3", 20274)]
    ref = synth.make_reference(sizes, seed=7, n_pad=200, n_families=6, family_len=120, family_copies=3,
                               n_long=2, long_len=400)
    reads = synth.make_reads(ref, 150, seed=8)
    run_case("case_basic", ref, reads,
             {"mam_l20": [], "mam_l16": ["-l", "16"], "mam_l25_n": ["-l", "25", "-n"], "mem_l20": ["-maxmatch"],
              "mem_l16": ["-maxmatch", "-l", "16"]}, tail=True)
    aref = adversarial_reference()
    areads = adversarial_reads(aref)
    run_case("case_adversarial", aref, areads,
             {"mam_l20": [], "mam_l12": ["-l", "12"], "mem_l20": ["-maxmatch"], "mem_l14": ["-maxmatch", "-l", "14"]}, tail=False)


if __name__ == "__main__":
    main()
