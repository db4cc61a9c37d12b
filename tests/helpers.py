"""Shared helpers of the test-suite (test infrastructure; may use oracle/)."""
import os

import numpy as np

from oracle import oracle as O
from oracle import tail as T
from smash_paper_b200 import synth


def make_case(path, **kw):
    """Tiny workload on disk + the reference-format index files written by the oracle's builder."""
    ref, reads, fa = synth.small_case(path, **kw)
    oix = O.Index.build(ref.names, ref.seqs)
    oix.save(fa)
    body = oix.mappability()
    with open(fa + ".bin/map.bin", "wb") as f:
        f.write(b"\x00\x00")
        f.write(body.tobytes())
    return ref, reads, fa, oix, body


def oracle_tail(oix, body, sam_bytes, case_dir, fa):
    lines = sam_bytes.splitlines(keepends=True)
    names = oix.descr[::2]
    sizes = oix.sizes[::2]
    tagged = T.tag_lines(lines, names, sizes, body)
    rows, nd, nn = T.smash_filter(tagged, names)
    pos = T.positions(rows)
    bins = T.read_table(os.path.join(case_dir, "bins.txt"))
    ci = T.read_chrominfo(fa + ".bin/chrom_sizes.txt")
    counts, total, dups, kept = T.varbin(pos, bins, ci)
    return dict(tagged=tagged, rows=rows, n_dupe=nd, n_non=nn, positions=pos, counts=np.array(counts),
                total=total, dups=dups, kept=kept, bins=bins, chrominfo=ci)
