"""Host-side mirror of the reference's query reader for `-samin` (QueryReader::run SAM branch,
query.cpp:639-648; NewQuery::extend query.cpp:125-144; Aligner::reset query.cpp:185-201) and of the
mapout writer (OutputSorter::flush, query.cpp:448-468).  Product code: numpy + bytes only."""
from __future__ import annotations

import os

import numpy as np

from .synth import ReadBatch


def _blob(items):
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        off[1:] = np.cumsum([len(x) for x in items])
    return np.frombuffer(b"".join(items), dtype=np.uint8).copy(), off


def derive_read_flag(name: bytes, flag: int):
    """`if (flag & is_first) name += ":0"; else if (flag & is_second) name += ":1";` followed by
    the strip of a trailing ':0' / ':1' into read_flag 65 / 129."""
    if flag & 64:
        name += b":0"
    elif flag & 128:
        name += b":1"
    rf = 0
    if len(name) >= 2 and name[-2:-1] == b":":
        if name[-1:] == b"0":
            name, rf = name[:-2], 65
        elif name[-1:] == b"1":
            name, rf = name[:-2], 129
    return name, rf


def parse_sam_lines(lines):
    """Whitespace-separated fields exactly like `input >> name >> flag >> ... >> errors` and then
    every further token as "\\t" + token.  Spaces inside SEQ are skipped (extend)."""
    names, seqs, quals, opts, flags, rfs = [], [], [], [], [], []
    for line in lines:
        line = line.rstrip(b"\n")
        if not line:
            continue
        f = line.split()
        if len(f) < 11:
            raise ValueError("SAM line with fewer than 11 fields: %r" % line[:60])
        try:
            flag = int(f[1])
        except ValueError:
            flag = 0
        name, rf = derive_read_flag(f[0], flag)
        names.append(name); flags.append(flag); rfs.append(rf)
        seqs.append(f[9]); quals.append(f[10])
        if len(f[10]) != len(f[9]):
            raise ValueError("SEQ and QUAL lengths differ for %r" % f[0])
        opts.append(b"".join(b"\t" + t for t in f[11:]))
    nb, no = _blob(names); sb, so = _blob(seqs); qb, _ = _blob(quals); ob, oo = _blob(opts)
    b = ReadBatch(names=nb, name_off=no, seq=sb, qual=qb, seq_off=so, flags=np.array(flags, dtype=np.uint16),
                  opt=ob, opt_off=oo)
    b.read_flag = np.array(rfs, dtype=np.uint16)
    return b


def read_sam(path):
    op = open
    if str(path).endswith(".gz"):
        import gzip
        op = gzip.open
    with op(path, "rb") as f:
        return parse_sam_lines(f.readlines())


def read_flag_of(batch):
    """read_flag array of a batch: parsed batches carry it; synthetic ones derive it from flags."""
    rf = getattr(batch, "read_flag", None)
    if rf is not None:
        return rf
    fl = batch.flags.astype(np.uint16)
    return np.where(fl & 64, 65, np.where(fl & 128, 129, 0)).astype(np.uint16)


def slice_batch(b, lo, hi):
    """Reads [lo,hi) of a batch (lo even keeps mates together)."""
    def cut(blob, off):
        return blob[off[lo]:off[hi]], (off[lo:hi + 1] - off[lo]).astype(np.int64)
    nb, no = cut(b.names, b.name_off); sb, so = cut(b.seq, b.seq_off); qb, _ = cut(b.qual, b.seq_off)
    ob, oo = cut(b.opt, b.opt_off) if b.opt.size else (b.opt, np.zeros(hi - lo + 1, dtype=np.int64))
    out = ReadBatch(names=np.ascontiguousarray(nb), name_off=no, seq=np.ascontiguousarray(sb), qual=np.ascontiguousarray(qb),
                    seq_off=so, flags=b.flags[lo:hi].copy(), opt=np.ascontiguousarray(ob), opt_off=oo)
    out.read_flag = np.ascontiguousarray(read_flag_of(b)[lo:hi])
    return out


def concat_batches(batches):
    def cat(blobs, offs):
        out_off = [np.zeros(1, dtype=np.int64)]
        base = 0
        for o in offs:
            out_off.append(o[1:] + base)
            base += int(o[-1])
        return np.concatenate(blobs), np.concatenate(out_off)
    nb, no = cat([b.names for b in batches], [b.name_off for b in batches])
    sb, so = cat([b.seq for b in batches], [b.seq_off for b in batches])
    qb, _ = cat([b.qual for b in batches], [b.seq_off for b in batches])
    ob, oo = cat([b.opt for b in batches], [b.opt_off for b in batches])
    out = ReadBatch(names=nb, name_off=no, seq=sb, qual=qb, seq_off=so, flags=np.concatenate([b.flags for b in batches]),
                    opt=ob, opt_off=oo)
    out.read_flag = np.concatenate([read_flag_of(b) for b in batches])
    return out


def write_mapout(header: bytes, sam: bytes, directory="mapout", tag="b200", seq=1):
    """mapout/mapout<id>.<k>.txt = header + records (query.cpp:453-463)."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, f"mapout{tag}.{seq}.txt")
    with open(path, "wb") as f:
        f.write(header)
        f.write(sam)
    return path
