"""`mummer`-compatible driver over libsmash_b200 (mummer.cpp:41-183): same single-dash flags, same
`<ref>.bin/` files, same `./mapout/*.txt` output, same exit codes and stderr shape.

    python -m smash_paper_b200.mummer -rcref -qthreads 12 -nomap -samin -samout ref.fa reads.sam
    python -m smash_paper_b200.mummer -verbose -rcref ref.fa dummy            # index build (index_setup.sh:19)
    python -m smash_paper_b200.mummer -rcref -mappability ref.fa ref.fa.bin/map.bin

Differences that cannot matter to the pipeline: chunk files are named mapoutb200.<k>.txt and hold
the records in input order (the reference's names embed a heap pointer and its chunk membership
depends on thread scheduling, query.cpp:453); -qthreads/-cached/-normalmem are accepted and ignored.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import api, samio, sequence

USAGE = ("Usage: {prog} [options] <reference-file> <query-file> ...\n"
         "-mum -mumreference -mumcand -maxmatch -l N -n -verbose -samin -samout -qthreads N -nomap -rcref\n"
         "-fastq -mappability -minblock N -cached -normalmem\n")
FLAGS0 = {"mumreference", "maxmatch", "mum", "mumcand", "n", "samout", "verbose", "nomap", "rcref", "fastq", "samin",
          "mappability", "cached", "normalmem"}
FLAGS1 = {"l", "qthreads", "minblock"}


class Args:
    def __init__(self, argv):
        self.min_len, self.mode, self.n, self.threads = 20, api.MODE_MAM, False, 2
        self.samout = self.verbose = self.nomap = self.rcref = self.fastq = self.samin = self.mappability = False
        pos = []
        i = 1
        while i < len(argv):
            a = argv[i]
            name = a.lstrip("-")
            if a.startswith("-") and name in FLAGS0:
                if name in ("mumreference", "mumcand"):
                    self.mode = api.MODE_MAM
                elif name == "maxmatch":
                    self.mode = api.MODE_MEM
                elif name == "mum":
                    self.mode = api.MODE_MUM
                elif name not in ("cached", "normalmem"):
                    setattr(self, name, True)
            elif a.startswith("-") and name in FLAGS1:
                i += 1
                v = int(argv[i])
                if name == "l":
                    self.min_len = v
                elif name == "qthreads":
                    self.threads = v
            elif a.startswith("-") and len(a) > 1:
                sys.stderr.write("Invalid arguments.\n" + USAGE.format(prog=argv[0]))
                raise SystemExit(1)
            else:
                pos.append(a)
            i += 1
        if len(pos) < 2:
            sys.stderr.write("There are too few arguments\n" + USAGE.format(prog=argv[0]))
            raise SystemExit(1)
        if self.fastq and self.samin:
            raise api.SmashError("-fastq cannot be used with -samin")
        if self.nomap and not self.samout:
            raise api.SmashError("-nomap can only be used with -sam_out")
        if self.mappability and not self.rcref:
            raise api.SmashError("-mappability requires -rcref")
        self.ref, self.inputs = pos[0], pos[1:]


def _index_exists(fa, rcref):
    base = f"{fa}.bin/rc{int(rcref)}"
    return any(os.access(f"{base}.i{w}.index.bin", os.R_OK) for w in (4, 8))


def _context(a: Args, need_isa):
    """Load `<ref>.bin/` (longSA load branch) or build it on the GPU and save it (build branch)."""
    if _index_exists(a.ref, a.rcref):
        if a.verbose:
            sys.stderr.write("# loading reference binary\n# loading index binary\n")
        ix = api.Index.open(a.ref, rcref=a.rcref)
        ctx = api.Context(ix, mode=a.mode, min_len=a.min_len, nomap=a.nomap, nucleotides_only=a.n)
        return ix, ctx
    if a.verbose:
        sys.stderr.write("# loading reference from fasta\n# creating index from reference\n")
    t0 = time.time()
    names, seqs = sequence.read_fasta(a.ref)
    text, startpos, sizes, descr = sequence.text_from_chromosomes(names, seqs, rcref=a.rcref)
    ctx = api.Context.from_text(text, startpos, sizes, descr, rcref=a.rcref, keep_isa=True, mode=a.mode,
                                min_len=a.min_len, nomap=a.nomap, nucleotides_only=a.n)
    if a.verbose:
        sys.stderr.write("# saving index\n")
    ctx.save_index(a.ref)
    if a.verbose:
        sys.stderr.write(f"# constructed index in {int(time.time() - t0)} seconds\n")
    ix = api.Index.open(a.ref, rcref=a.rcref)
    return ix, ctx


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    try:
        a = Args(argv)
        ix, ctx = _context(a, need_isa=a.mappability)
        if a.mappability:
            hdr = ix.sam_header().decode().splitlines()
            total = sum(int(l.split("LN:")[1]) for l in hdr if l.startswith("@SQ"))
            body = ctx.build_mappability(total)
            with open(a.inputs[0], "wb") as f:
                f.write(b"\x00\x00")                      # the reference's two junk bytes (longSA.cpp:606-617)
                f.write(body.tobytes())
            return 0
        header = ix.sam_header()
        n_q = 0
        t0 = time.time()
        if a.verbose:
            sys.stderr.write(f"# running {a.threads} threads to answer queries\n# running {len(a.inputs)} query reader\n")
        for k, path in enumerate(a.inputs):
            if not os.access(path, os.R_OK):
                sys.stderr.write(f"unable to open {path}\n")
                return 1
            if not a.samin:
                raise api.SmashError("only -samin input is implemented by the GPU driver")
            batch = samio.read_sam(path)
            n_q += batch.n
            step = 1 << 20
            chunk = 0
            for lo in range(0, batch.n, step):
                hi = min(batch.n, lo + step)
                sub = samio.slice_batch(batch, lo, hi)
                # a chunk's lines in OutputSorter order (query.cpp:448-468) unless SMASH_CHUNK_ORDER=input
                want = api.WANT_SAM | (0 if os.environ.get("SMASH_CHUNK_ORDER") == "input" else api.WANT_SORTED)
                res = ctx.map_batch(sub, want=want)
                if a.samout:
                    chunk += 1
                    samio.write_mapout(header, res.sam, tag=f"b200_{k}", seq=chunk)
        if a.verbose:
            sys.stderr.write(f"# ran {n_q} queries in {int(time.time() - t0)} seconds\n")
        return 0
    except api.SmashError as e:
        sys.stderr.write("Error\n" + str(e) + "\n")
        return 1


if __name__ == "__main__":
    raise SystemExit(main())
