"""Host-side mirror of the reference's `Sequence` (fasta.h:24-42, fasta.cpp:89-252): FASTA ->
one lower-case text `chr_fwd ` chr_rc ` ... $` plus startpos/sizes/descr.  Product code (used by
the mummer-compatible driver and bench.py to feed smash_ctx_create_from_text); numpy only."""
from __future__ import annotations

import numpy as np

_COMP = np.arange(256, dtype=np.uint8)                      # reverse_complement, fasta.cpp:26-61
for _a, _b in zip(b"acgtrymkbdhvACGTRYMKBDHV", b"tgcayrkmvhdbTGCAYRKMVHDB"):
    _COMP[_a] = _b
_LOWER = np.arange(256, dtype=np.uint8)
_LOWER[65:91] += 32


def text_from_chromosomes(names, seqs, rcref=True):
    """seqs: uint8 arrays (any case).  Returns (text, startpos, sizes, descr) exactly as the
    FASTA branch of Sequence::Sequence lays them out (fasta.cpp:151-203)."""
    n = len(names)
    total = sum(len(s) for s in seqs)
    N = (2 * total + 2 * n) if rcref else (total + n)
    text = np.empty(N, dtype=np.uint8)
    startpos, sizes, descr = [], [], []
    pos = 0
    for k, (name, s) in enumerate(zip(names, seqs)):
        last = k == n - 1
        L = len(s)
        low = _LOWER[s]
        startpos.append(pos); sizes.append(L); descr.append(name)
        text[pos:pos + L] = low
        pos += L
        if rcref or not last:
            text[pos] = 0x60
            pos += 1
        if rcref:
            startpos.append(pos); sizes.append(L); descr.append(name)
            text[pos:pos + L] = _COMP[low[::-1]]
            pos += L
            if not last:
                text[pos] = 0x60
                pos += 1
    text[pos] = 0x24
    pos += 1
    assert pos == N
    return text, startpos, sizes, descr


def read_fasta(path):
    """Name = header up to the first space (fasta.cpp:192-195); sequence lines concatenated."""
    names, seqs, cur = [], [], []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if not line:
                continue
            if line.startswith(b">"):
                if cur or names:
                    seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
                    cur = []
                names.append(line[1:].strip(b" ").split(b" ")[0].decode())
            else:
                cur.append(line.strip(b" "))
    if names:
        seqs.append(np.frombuffer(b"".join(cur), dtype=np.uint8))
    return names, seqs
