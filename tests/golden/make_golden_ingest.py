#!/usr/bin/env python
"""Golden fixtures of the INPUT side, produced by the UNMODIFIED reference binaries
(oracle/_ref/fastqs_to_sam, oracle/_ref/mummer; built from /root/reference by oracle/Makefile).  Run in the dev
container only:

    python tests/golden/make_golden_ingest.py

tests/golden/case_ingest/ (reference text = case_basic/ref.fa.gz):
  r1.fq.gz, r2.fq.gz        FASTQ mate files with the quirks fastqs_to_sam.cpp:47-95 handles: second header token,
                            blank lines before '@' and '+', CRLF records, '>' records, lower-case n, an empty read
                            in one mate only (shifts the pairing by arrival parity from there on), unequal record
                            counts, no newline at the end
  fastqs_to_sam_{0,1}.sam.gz  what `fastqs_to_sam r1.fq r2.fq [1]` printed
  mapout_fastq.sam.gz       `mummer -rcref -nomap -samin -samout` over fastqs_to_sam_1.sam (smash_mapping.sh:19)
  quirks.sam.gz             SAM lines exercising `istringstream >>` (query.cpp:640-648): runs of blanks and tabs
                            as separators, CRLF, "77x" (the x becomes the next field), +77 / 0077 / -1 flags, names
                            ending in :0 / :1, several optional fields, empty lines
  mapout_quirks.sam.gz      the reference's records for it
"""
import gzip
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
from smash_paper_b200 import sequence, synth  # noqa: E402
from make_golden import gz_write  # noqa: E402


def fastq_pair(reads, rng):
    """Interleaved ReadBatch (mate 1 at even, mate 2 at odd index) -> two FASTQ texts with quirks."""
    out = [[], []]
    n_pairs = reads.n // 2
    for k in range(n_pairs):
        for f in (0, 1):
            i = 2 * k + f
            name = bytes(reads.names[reads.name_off[i]:reads.name_off[i + 1]])
            seq = bytes(reads.seq[reads.seq_off[i]:reads.seq_off[i + 1]]).replace(b"Z", b"N")
            qual = bytes(reads.qual[reads.seq_off[i]:reads.seq_off[i + 1]])
            eol = b"\r\n" if k % 17 == 3 else b"\n"
            hdr = b"@" + name
            if k % 5 == 1:
                hdr += b" %d:N:0:ACGT extra words" % (f + 1)
            if k % 11 == 4:
                hdr = b"  " + hdr + b"\t"
            pre = b"\n\n" if k % 13 == 6 else b""
            mid = b"\n \n" if k % 19 == 7 else b""
            if k == 40 and f == 1:
                rec = b">" + name + b" fasta_record" + eol + seq + eol              # '>' record: errors = bases
            elif k == n_pairs - 20 and f == 0:
                rec = hdr + eol + eol + b"+" + eol + eol                            # empty read: prints nothing
            elif k == 25:
                rec = hdr + eol + seq[:70] + b"nn" + seq[72:] + eol + b"+" + name + eol + qual + eol
            else:
                rec = pre + hdr + eol + seq + eol + mid + b"+" + eol + qual + eol
            out[f].append(rec)
    a, b = b"".join(out[0]), b"".join(out[1][:-3])                                  # mate 2 is three records short
    return a[:-1], b                                                                  # mate 1 without a final newline


def quirks_sam(reads):
    lines = []
    for i in range(0, 60):
        name = bytes(reads.names[reads.name_off[i]:reads.name_off[i + 1]])
        seq = bytes(reads.seq[reads.seq_off[i]:reads.seq_off[i + 1]])
        qual = bytes(reads.qual[reads.seq_off[i]:reads.seq_off[i + 1]])
        flag = b"77" if i % 2 == 0 else b"141"
        mid = [b"*", b"0", b"0", b"*", b"*", b"0", b"0"]
        sep, eol, opt = b"\t", b"\n", b""
        k = i // 2
        if k == 1:
            sep = b"  \t "
        elif k == 2:
            eol = b"\r\n"
        elif k == 3:
            flag += b"x"; mid = mid[1:]                        # "77x": the x is read as the next field
        elif k == 4:
            flag = b"+" + flag
        elif k == 5:
            flag = b"00" + flag
        elif k == 6:
            flag = b"-1"                                       # unsigned wrap: bit 64 set -> ":0" for both reads
        elif k == 7:
            flag = b"0"; name += b":0" if i % 2 == 0 else b":1"
        elif k == 8:
            flag = b"4"; name += b":1" if i % 2 == 0 else b":7"
        elif k == 9:
            opt = b"\tXA:Z:first  \t YB:i:22\tZC:Z:last \t"
        elif k == 10:
            opt = b" NM:i:0"
            eol = b" \r\n"
        elif k == 11:
            name = b" \t" + name
        elif k == 12:
            name += b":0"                                      # with flag 77 / 141: suffix appended, then one suffix cut
        line = sep.join([name, flag] + mid + [seq, qual]) + opt + eol
        lines.append(line)
        if k in (13, 14) and i % 2 == 1:
            lines.append(b"\n\n")
    # the text ends with a newline: without one the reference's `while (data) { getline(data, line); if (line.size())`
    # (query.cpp:625-627) sees the stale last line a second time and OutputSorter then throws "flags equal"
    return b"".join(lines)


def main():
    assert O.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    out = os.path.join(HERE, "case_ingest")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "ref.fa")
        open(fa, "wb").write(gzip.open(os.path.join(HERE, "case_basic", "ref.fa.gz")).read())
        names, seqs = sequence.read_fasta(fa)
        ref = synth.Reference(names, [np.asarray(s, dtype=np.uint8) for s in seqs])
        O.ref_build_index(fa, mappability=False)
        reads = synth.make_reads(ref, 120, seed=77)
        fq1, fq2 = fastq_pair(reads, rng)
        open(os.path.join(d, "r1.fq"), "wb").write(fq1)
        open(os.path.join(d, "r2.fq"), "wb").write(fq2)
        gz_write(os.path.join(out, "r1.fq.gz"), fq1)
        gz_write(os.path.join(out, "r2.fq.gz"), fq2)
        exe = os.path.join(O.REF_BIN, "fastqs_to_sam")
        for rep in (0, 1):
            args = [exe, os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")] + (["1"] if rep else [])
            sam = subprocess.run(args, check=True, stdout=subprocess.PIPE).stdout
            gz_write(os.path.join(out, f"fastqs_to_sam_{rep}.sam.gz"), sam)
            print("fastqs_to_sam", rep, sam.count(b"\n"), "lines")
        open(os.path.join(d, "fq.sam"), "wb").write(sam)
        hdr, lines = O.ref_map(fa, os.path.join(d, "fq.sam"), d)
        gz_write(os.path.join(out, "mapout_fastq.sam.gz"), hdr + b"".join(lines))
        print("mapout_fastq", len(lines), "records")
        q = quirks_sam(synth.make_reads(ref, 40, seed=78))
        open(os.path.join(d, "quirks.sam"), "wb").write(q)
        gz_write(os.path.join(out, "quirks.sam.gz"), q)
        hdr, lines = O.ref_map(fa, os.path.join(d, "quirks.sam"), d)
        gz_write(os.path.join(out, "mapout_quirks.sam.gz"), hdr + b"".join(lines))
        print("mapout_quirks", len(lines), "records")


if __name__ == "__main__":
    main()
