"""Deterministic synthetic inputs for the SMASH hot path.

Everything here is *test/bench input generation*: a random reference with planted repeats and
N-padded chromosome ends, chimeric SMASH-style read pairs, bin tables and the small text files
`index_setup.sh` derives from `samtools faidx` (chrom_sizes.txt, sam_header.txt).  All draws
come from `numpy.random.default_rng(seed)` so the same arguments give the same bytes on every
machine.  Nothing in this module is on the product path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# hg19 chr1-22,X,Y lengths, recovered from sample_bins/50000/bins.txt (last stop_chrpos per
# chromosome); order matters: abs offsets must equal bins.txt column 3.
HG19_SIZES = [
    ("chr1", 249250621), ("chr2", 243199373), ("chr3", 198022430), ("chr4", 191154276),
    ("chr5", 180915260), ("chr6", 171115067), ("chr7", 159138663), ("chr8", 146364022),
    ("chr9", 141213431), ("chr10", 135534747), ("chr11", 135006516), ("chr12", 133851895),
    ("chr13", 115169878), ("chr14", 107349540), ("chr15", 102531392), ("chr16", 90354753),
    ("chr17", 81195210), ("chr18", 78077248), ("chr19", 59128983), ("chr20", 63025520),
    ("chr21", 48129895), ("chr22", 51304566), ("chrX", 155270560), ("chrY", 59373566),
]

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_a] = _b


@dataclass
class Reference:
    names: list
    seqs: list  # list of uint8 arrays (upper-case ASCII)
    sizes: list = field(default_factory=list)

    def __post_init__(self):
        self.sizes = [int(len(s)) for s in self.seqs]

    @property
    def total(self):
        return int(sum(self.sizes))

    def offsets(self):
        off, out = 0, []
        for s in self.sizes:
            out.append(off)
            off += s
        return out

    def concat(self):
        return np.concatenate(self.seqs)


def make_reference(chrom_sizes, seed=1, n_pad=1000, n_families=200, family_len=300,
                   family_copies=5, n_long=4, long_len=600, n_highcopy=0, highcopy_len=64,
                   highcopy_copies=4200):
    """Random ACGT chromosomes with planted repeat families (some reverse-complemented), a few
    >255 bp repeats (LCP overflow table) and optional high-copy families (to trip the
    reference's expand_link threshold, longSA.h:159); `n_pad` N's at both ends of every
    chromosome (keeps mappability_tag from throwing at chromosome ends)."""
    rng = np.random.default_rng(seed)
    names = [n for n, _ in chrom_sizes]
    seqs = []
    for _, size in chrom_sizes:
        # 4 bases per random byte (one PCG64 byte stream per chromosome)
        raw = np.frombuffer(rng.bytes((size + 3) // 4), dtype=np.uint8)
        s = np.empty(4 * len(raw), dtype=np.uint8)
        for k in range(4):
            s[k::4] = _ACGT[(raw >> (2 * k)) & 3]
        seqs.append(s[:size].copy() if 4 * len(raw) != size else s)
    lens = np.array([len(s) for s in seqs])

    def plant(unit, copies):
        L = len(unit)
        for _ in range(copies):
            c = int(rng.integers(0, len(seqs)))
            if lens[c] < 2 * n_pad + L + 2:
                continue
            p = int(rng.integers(n_pad, lens[c] - n_pad - L))
            u = unit
            if rng.random() < 0.3:
                u = _COMP[unit[::-1]]
            seqs[c][p:p + L] = u

    for _ in range(n_families):
        plant(_ACGT[rng.integers(0, 4, size=family_len, dtype=np.uint8)], family_copies)
    for _ in range(n_long):
        plant(_ACGT[rng.integers(0, 4, size=long_len, dtype=np.uint8)], 3)
    for _ in range(n_highcopy):
        plant(_ACGT[rng.integers(0, 4, size=highcopy_len, dtype=np.uint8)], highcopy_copies)
    for s in seqs:
        if n_pad and len(s) > 2 * n_pad:
            s[:n_pad] = ord("N")
            s[-n_pad:] = ord("N")
    return Reference(names, seqs)


def write_fasta(ref: Reference, path, width=60):
    with open(path, "wb") as f:
        for name, s in zip(ref.names, ref.seqs):
            f.write(b">" + name.encode() + b"\n")
            n = len(s)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def write_index_side_files(ref: Reference, fasta_path):
    """chrom_sizes.txt / sam_header.txt as index_setup.sh:28,31 derive them from the .fai."""
    d = fasta_path + ".bin"
    os.makedirs(d, exist_ok=True)
    off = 0
    with open(os.path.join(d, "chrom_sizes.txt"), "w") as f:
        for n, s in zip(ref.names, ref.sizes):
            if "_" in n:
                continue
            f.write(f"{n}\t{s}\t{off}\n")
            off += s
    with open(os.path.join(d, "sam_header.txt"), "w") as f:
        for n, s in zip(ref.names, ref.sizes):
            f.write(f"@SQ\tSN:{n}\tLN:{s}\n")


def write_fixed_bins(ref: Reference, path, width=50000):
    """bins.txt: chr start_chrpos start_abspos stop_chrpos bin_len n_maps_expected (binning.sh:22-24)."""
    rows = []
    off = 0
    for n, s in zip(ref.names, ref.sizes):
        for st in range(0, s, width):
            en = min(st + width, s)
            rows.append(f"{n}\t{st}\t{off + st}\t{en}\t{en - st}\t{en - st}\n")
        off += s
    with open(path, "w") as f:
        f.writelines(rows)
    return len(rows)


def split_bins(src, dst, parts):
    """Synthesise finer bins by splitting every row of `src` into `parts` equal pieces
    (sample_bins/100000 and /500000 ship without bins.txt)."""
    out = []
    for line in open(src):
        c, st, ab, en, ln, ex = line.rstrip("\n").split("\t")
        st, ab, en = int(st), int(ab), int(en)
        n = en - st
        for k in range(parts):
            a = st + (n * k) // parts
            b = st + (n * (k + 1)) // parts
            if b > a:
                out.append(f"{c}\t{a}\t{ab + a - st}\t{b}\t{b - a}\t{b - a}\n")
    with open(dst, "w") as f:
        f.writelines(out)
    return len(out)


@dataclass
class ReadBatch:
    """Packed read batch = what crosses the C-ABI (include/smash_b200.h: smash_batch_t)."""
    names: np.ndarray       # uint8 blob (no :0/:1 suffix)
    name_off: np.ndarray    # int64[n+1]
    seq: np.ndarray         # uint8 blob, original case
    qual: np.ndarray        # uint8 blob
    seq_off: np.ndarray     # int64[n+1]
    flags: np.ndarray       # uint16[n] input SAM flag (77/141)
    opt: np.ndarray         # uint8 blob: "\tTAG..." per read (may be empty)
    opt_off: np.ndarray     # int64[n+1]

    @property
    def n(self):
        return len(self.flags)


def make_reads(ref: Reference, n_pairs, read_len=150, seed=2, frag_min=3, frag_max=8,
               sub_rate=0.0075, z_rate=0.01, random_frac=0.02, dup_frac=0.002,
               first_pair=0, concat=None, chunk=100000):
    """Chimeric SMASH-like read pairs: every read is a concatenation of `frag_min..frag_max`
    fragments from uniform random loci/strands of `ref`, with substitutions, 'Z' bases
    (fastqs_to_sam.cpp:69 N->Z), a few fully random reads and a few exact duplicate pairs.
    Names are r%09d so byte order == numeric order == input order."""
    n = 2 * n_pairs
    q = read_len
    genome = ref.concat() if concat is None else concat
    G = len(genome)
    seq = np.empty((n, q), dtype=np.uint8)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = c1 - c0
        rng = np.random.default_rng([seed, first_pair, c0])
        nfrag = rng.integers(frag_min, frag_max + 1, size=m)
        cuts = np.sort(rng.integers(1, q, size=(m, frag_max - 1)), axis=1)
        # keep only the first nfrag-1 cuts (others pushed past the read end)
        cuts = np.where(np.arange(frag_max - 1)[None, :] < (nfrag - 1)[:, None], cuts, q)
        j = np.arange(q)
        fid = (cuts[:, None, :] <= j[None, :, None]).sum(axis=2)            # (m,q) fragment id
        starts = np.concatenate([np.zeros((m, 1), dtype=np.int64), cuts], axis=1)  # (m,fmax)
        ends = np.concatenate([cuts, np.full((m, 1), q)], axis=1)
        locus = rng.integers(0, G - q, size=(m, frag_max))
        strand = rng.integers(0, 2, size=(m, frag_max)).astype(bool)
        fs = np.take_along_axis(starts, fid, axis=1)
        fe = np.take_along_axis(ends, fid, axis=1)
        lo = np.take_along_axis(locus, fid, axis=1)
        st = np.take_along_axis(strand, fid, axis=1)
        idx = np.where(st, lo + (fe - 1 - j[None, :]), lo + (j[None, :] - fs))
        b = genome[idx]
        b = np.where(st, _COMP[b], b)
        sub = rng.random((m, q)) < sub_rate
        b = np.where(sub, _ACGT[rng.integers(0, 4, size=(m, q))], b)
        rnd = rng.random(m) < random_frac
        if rnd.any():
            b[rnd] = _ACGT[rng.integers(0, 4, size=(int(rnd.sum()), q))]
        z = rng.random((m, q)) < z_rate
        b = np.where(z | (b == ord("N")), ord("Z"), b)
        seq[c0:c1] = b
    rng = np.random.default_rng([seed, first_pair, 12345])
    ndup = int(n_pairs * dup_frac)
    if ndup and n_pairs > 2:
        src = rng.integers(0, n_pairs - 1, size=ndup)
        dst = np.minimum(src + rng.integers(1, 50, size=ndup), n_pairs - 1)
        for s, d in zip(src, dst):
            if s != d:
                seq[2 * d] = seq[2 * s]
                seq[2 * d + 1] = seq[2 * s + 1]
    qual = (33 + 2 + rng.integers(0, 39, size=(n, q))).astype(np.uint8)
    ids = first_pair + np.arange(n_pairs)
    name_arr = np.char.add("r", np.char.zfill(ids.astype(str), 9)).astype("S10")
    name_blob = np.repeat(np.frombuffer(name_arr.tobytes(), dtype=np.uint8).reshape(n_pairs, 10),
                          2, axis=0)
    flags = np.tile(np.array([77, 141], dtype=np.uint16), n_pairs)
    return ReadBatch(
        names=name_blob.reshape(-1).copy(),
        name_off=np.arange(n + 1, dtype=np.int64) * 10,
        seq=seq.reshape(-1), qual=qual.reshape(-1),
        seq_off=np.arange(n + 1, dtype=np.int64) * q,
        flags=flags, opt=np.zeros(0, dtype=np.uint8), opt_off=np.zeros(n + 1, dtype=np.int64))


_FAST = None


def _fast_lib():
    """smash_paper_b200/host/synth_reads.c compiled on first use (gcc, in-tree .so)."""
    global _FAST
    if _FAST is None:
        import ctypes
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        src = os.path.join(here, "host", "synth_reads.c")
        so = os.path.join(here, "host", "libsynth_reads.so")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-pthread", "-o", so, src])
        _FAST = ctypes.CDLL(so)
    return _FAST


def make_reads_fast(genome, n_pairs, read_len=150, seed=2, frag_min=3, frag_max=8, sub_rate=0.0075,
                    z_rate=0.01, random_frac=0.02, dup_frac=0.002, first_pair=0, n_threads=None):
    """Same recipe as make_reads, generated by host/synth_reads.c (~100x faster; a different but
    equally deterministic random stream: every read is a function of (seed, global pair index))."""
    import ctypes as C
    n = 2 * n_pairs
    q = read_len
    seq = np.empty(n * q, dtype=np.uint8)
    qual = np.empty(n * q, dtype=np.uint8)
    genome = np.ascontiguousarray(genome, dtype=np.uint8)
    _fast_lib().synth_reads(genome.ctypes.data_as(C.c_void_p), C.c_uint64(len(genome)), C.c_uint64(seed),
                            C.c_uint64(first_pair), C.c_uint64(n_pairs), C.c_int(q), C.c_int(frag_min), C.c_int(frag_max),
                            C.c_double(sub_rate), C.c_double(z_rate), C.c_double(random_frac), C.c_double(dup_frac),
                            seq.ctypes.data_as(C.c_void_p), qual.ctypes.data_as(C.c_void_p),
                            C.c_int(n_threads or min(16, os.cpu_count() or 1)))
    ids = first_pair + np.arange(n_pairs, dtype=np.int64)
    digits = np.empty((n_pairs, 10), dtype=np.uint8)
    digits[:, 0] = ord("r")
    v = ids.copy()
    for k in range(9, 0, -1):
        digits[:, k] = 48 + (v % 10)
        v //= 10
    name_blob = np.repeat(digits, 2, axis=0)
    return ReadBatch(names=name_blob.reshape(-1), name_off=np.arange(n + 1, dtype=np.int64) * 10, seq=seq, qual=qual,
                     seq_off=np.arange(n + 1, dtype=np.int64) * q, flags=np.tile(np.array([77, 141], dtype=np.uint16), n_pairs),
                     opt=np.zeros(0, dtype=np.uint8), opt_off=np.zeros(n + 1, dtype=np.int64))


def sam_text_fast(batch: ReadBatch):
    """The text write_sam() writes, as one uint8 array, built with array operations.  Only for batches of the
    make_reads_fast shape: names and reads of one fixed length each, flags 77/141 alternating, no optional fields."""
    n = batch.n
    assert n % 2 == 0 and n > 0 and batch.opt.size == 0
    ln, q = int(batch.name_off[1]), int(batch.seq_off[1])
    assert batch.names.size == n * ln and batch.seq.size == n * q
    mid = [np.frombuffer(b"\t77\t*\t0\t0\t*\t*\t0\t0\t", dtype=np.uint8), np.frombuffer(b"\t141\t*\t0\t0\t*\t*\t0\t0\t", dtype=np.uint8)]
    names = batch.names.reshape(n // 2, 2, ln); seq = batch.seq.reshape(n // 2, 2, q); qual = batch.qual.reshape(n // 2, 2, q)
    cols = []
    for m in (0, 1):
        w = ln + len(mid[m]) + q + 1 + q + 1
        row = np.empty((n // 2, w), dtype=np.uint8)
        o = 0
        row[:, o:o + ln] = names[:, m]; o += ln
        row[:, o:o + len(mid[m])] = mid[m]; o += len(mid[m])
        row[:, o:o + q] = seq[:, m]; o += q
        row[:, o] = 9; o += 1
        row[:, o:o + q] = qual[:, m]; o += q
        row[:, o] = 10
        cols.append(row)
    return np.concatenate(cols, axis=1).reshape(-1)


def write_sam(batch: ReadBatch, path):
    """Unaligned SAM lines as fastqs_to_sam.cpp:80-93 prints them (flags 77/141)."""
    with open(path, "wb") as f:
        nb, sb, qb, ob = (batch.names.tobytes(), batch.seq.tobytes(), batch.qual.tobytes(),
                          batch.opt.tobytes())
        no, so, oo = batch.name_off, batch.seq_off, batch.opt_off
        out = []
        for i in range(batch.n):
            out.append(nb[no[i]:no[i + 1]] + b"\t%d\t*\t0\t0\t*\t*\t0\t0\t" % batch.flags[i]
                       + sb[so[i]:so[i + 1]] + b"\t" + qb[so[i]:so[i + 1]]
                       + ob[oo[i]:oo[i + 1]] + b"\n")
            if len(out) >= 65536:
                f.write(b"".join(out))
                out = []
        f.write(b"".join(out))


def small_case(tmpdir, n_chrom=3, chrom_len=20000, n_pairs=200, seed=7, read_len=150, **kw):
    """A complete tiny workload on disk: ref.fa (+side files), reads.sam, bins.txt."""
    os.makedirs(tmpdir, exist_ok=True)
    sizes = [(f"chr{i + 1}", chrom_len + 137 * i) for i in range(n_chrom)]
    ref = make_reference(sizes, seed=seed, n_pad=kw.pop("n_pad", 200),
                         n_families=kw.pop("n_families", 6), family_len=kw.pop("family_len", 120),
                         family_copies=3, n_long=kw.pop("n_long", 2), long_len=400,
                         n_highcopy=kw.pop("n_highcopy", 0), highcopy_len=40,
                         highcopy_copies=kw.pop("highcopy_copies", 300))
    fa = os.path.join(tmpdir, "ref.fa")
    write_fasta(ref, fa)
    write_index_side_files(ref, fa)
    reads = make_reads(ref, n_pairs, read_len=read_len, seed=seed + 1, **kw)
    write_sam(reads, os.path.join(tmpdir, "reads.sam"))
    write_fixed_bins(ref, os.path.join(tmpdir, "bins.txt"), width=5000)
    return ref, reads, fa
