"""oracle/tail.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python restatement of the tail of smash_mapping.sh / binning.sh:

  mappability_tag  (mappability_tag.cpp:53-128, util.h:131-149, chromosomes.h:29-111)  PINNED:
                   checked against the compiled reference binary (tests/test_oracle_vs_reference.py)
  smashMEM.py      (smashMEM.py:84-92, 154-230 with argv "0 0 10000 4")            PARITY UNPINNED:
                   pysam + samtools are not installable here and the reference ships no vector for
                   this stage; restated from the source and pysam 0.7-era accessor semantics
                   (qstart = leading soft clip, qend = l_qseq - trailing soft clip, qlen = qend-qstart,
                   rlen = l_qseq, pos = POS-1, tid = @SQ index).
  awk/perl filter  (smash_mapping.sh:29)
  varbin.py        (varbin.py:6-118)                                                PINNED: checked
                   against the reference script run under python3 (cols 1-5; stats via '//').
"""
from __future__ import annotations

import bisect
import re

import numpy as np

_CIG = re.compile(rb"(\d+)([A-Za-z=])")


def chrom_offsets32(names, sizes):
    """ChromosomeInfo(sam_header.txt) offsets: 32-bit running sums over ALL @SQ lines."""
    off, acc = {}, 0
    for n, s in zip(names, sizes):
        off[n] = acc & 0xFFFFFFFF
        acc = (acc + int(s)) & 0xFFFFFFFF
    return off


def tag_line(line: bytes, offsets, mapbody: np.ndarray):
    """One SAM record -> record + L<i>/R<i> tags.  `mapbody` = map.bin without its 2 junk bytes.
    Raises ValueError where the reference throws (mappability_tag.cpp:107-113, 119)."""
    core = line.rstrip(b"\n")
    f = core.split()
    chrom, pos, cigar = f[2].decode(), int(f[3]) & 0xFFFFFFFF, f[5]
    small = "_gl000" in chrom or "chrM" in chrom
    opt = b""
    if cigar != b"*":
        abspos = (offsets[chrom] + pos) & 0xFFFFFFFF
        off, u = 0, 0
        for cnt, code in _CIG.findall(cigar):
            cnt = int(cnt)
            if code == b"=":
                li = (abspos + off + cnt - 1) & 0xFFFFFFFF
                ri = (abspos + off - 1) & 0xFFFFFFFF
                lm = int(mapbody[2 * li]) if 2 * li < len(mapbody) else 0
                rm = int(mapbody[2 * ri + 1]) if 2 * ri + 1 < len(mapbody) else 0
                left = lm - 1 if lm else 255
                right = rm if rm else 255
                if u < 10:
                    opt += b"\tL%d:i:%d\tR%d:i:%d" % (u, left, u, right)
                if left > cnt and not small:
                    raise ValueError("left mappability too big %d" % left)
                if right > cnt and not small:
                    raise ValueError("right mappability too big %d" % right)
                u += 1
            elif code not in (b"S", b"M"):
                raise ValueError("unexpected cigar %r" % code)
            off += cnt
    return core + opt + b"\n"


def tag_lines(lines, names, sizes, mapbody):
    offs = chrom_offsets32(names, sizes)
    return [ln if ln.startswith(b"@") else tag_line(ln, offs, mapbody) for ln in lines]


def strnum_cmp(a: bytes, b: bytes) -> int:
    """`samtools sort -n` name order (smash_mapping.sh:23).  samtools is a third-party dependency that is absent from
    /root/reference and from this image: the legacy `sort -n - prefix` syntax the script uses is samtools 0.1.x, whose
    bam_sort.c compares query names with strnum_cmp -- digit runs compare as numbers (leading zeros skipped; equal
    numbers: the run with FEWER leading zeros is greater), everything else bytewise -- and breaks ties with
    flag & 0xc0 (read 1 before read 2) in a stable merge sort.  Restated here from that published algorithm."""
    na, nb = len(a), len(b)
    pa = pb = 0

    def dig(s, i):
        return i < len(s) and 48 <= s[i] <= 57

    while pa < na and pb < nb:
        if dig(a, pa) and dig(b, pb):
            while pa < na and a[pa] == 48:
                pa += 1
            while pb < nb and b[pb] == 48:
                pb += 1
            while dig(a, pa) and dig(b, pb) and a[pa] == b[pb]:
                pa += 1
                pb += 1
            if dig(a, pa) and dig(b, pb):
                i = 0
                while dig(a, pa + i) and dig(b, pb + i):
                    i += 1
                return 1 if dig(a, pa + i) else -1 if dig(b, pb + i) else a[pa] - b[pb]
            if dig(a, pa):
                return 1
            if dig(b, pb):
                return -1
            if pa != pb:
                return 1 if pa < pb else -1
        else:
            if a[pa] != b[pb]:
                return a[pa] - b[pb]
            pa += 1
            pb += 1
    return 1 if pa < na else -1 if pb < nb else 0


def name_sort_lines(record_lines):
    """Record lines (no header) in `samtools sort -n` order: strnum_cmp(qname), then flag & 0xc0, stable."""
    import functools

    def key(ln):
        f = ln.split(b"\t", 2)
        return f[0], int(f[1]) & 0xC0

    def cmp(x, y):
        return strnum_cmp(x[0][0], y[0][0]) or (x[0][1] - y[0][1])

    keyed = [(key(ln), ln) for ln in record_lines]
    keyed.sort(key=functools.cmp_to_key(cmp))
    return [ln for _, ln in keyed]


class _Hit:
    __slots__ = ("name", "flag", "tid", "chrom", "pos", "rev", "rlen", "qstart", "qend", "qlen",
                 "L0", "R0", "HI", "unmapped", "read1", "read2")


def _parse_hit(line: bytes, tid_of):
    f = line.rstrip(b"\n").split(b"\t")
    h = _Hit()
    h.name = f[0]
    h.flag = int(f[1])
    h.unmapped = bool(h.flag & 4)
    h.read1 = bool(h.flag & 64)
    h.read2 = bool(h.flag & 128)
    h.rev = bool(h.flag & 16)
    h.chrom = f[2].decode()
    h.tid = tid_of.get(h.chrom, -1)
    h.pos = int(f[3]) - 1
    h.rlen = len(f[9])
    ops = [(int(c), o) for c, o in _CIG.findall(f[5])] if f[5] != b"*" else []
    lead = 0
    for c, o in ops:
        if o == b"S":
            lead += c
        elif o != b"H":
            break
    trail = 0
    for c, o in reversed(ops):
        if o == b"S":
            trail += c
        elif o != b"H":
            break
    h.qstart, h.qend = lead, h.rlen - trail
    h.qlen = h.qend - h.qstart
    h.L0 = h.R0 = h.HI = None
    for t in f[11:]:
        if t.startswith(b"L0:i:"):
            h.L0 = int(t[5:])
        elif t.startswith(b"R0:i:"):
            h.R0 = int(t[5:])
        elif t.startswith(b"HI:i:"):
            h.HI = int(t[5:])
    return h


def smash_filter(tagged_lines, chrom_names, min_match=0, min_ratio=0.0, hit_window=10000,
                 min_excess=4, presorted=False):
    """smashMEM.py main loop.  Returns (rows, n_dupe, n_non_dupe); rows are the printed tuples
    (readID, read_index, hit_index, chrom, pos, reverse, read_len, hit_offset, match_len, umatch,
    excess) in output order."""
    tid_of = {n: i for i, n in enumerate(chrom_names)}
    hits = [_parse_hit(ln, tid_of) for ln in tagged_lines if not ln.startswith(b"@")]
    if not presorted:
        import functools
        hits.sort(key=functools.cmp_to_key(lambda x, y: strnum_cmp(x.name, y.name) or ((x.flag & 0xC0) - (y.flag & 0xC0))))
    rows, dupes, n_dupe, n_non = [], set(), 0, 0
    i = 0
    while i < len(hits):
        j = i
        while j < len(hits) and hits[j].name == hits[i].name:
            j += 1
        group = hits[i:j]
        i = j
        r1 = [h for h in group if h.read1]
        r2 = [h for h in group if not h.read1]
        if not r1:
            raise IndexError("reads1[0]: read without read-1 records (smashMEM.py:156)")
        rid = r1[0].name

        def excess_ok(hs):                                       # smashMEM.py:84-92
            return [h for h in hs if not h.unmapped and h.qlen - max(h.L0, h.R0) >= min_excess]

        def match_ok(hs):                                        # smashMEM.py:74-80
            return [h for h in hs if not h.unmapped and h.qlen >= min_match]

        r1, r2 = match_ok(excess_ok(r1)), match_ok(excess_ok(r2))
        if not r1 and not r2:
            continue

        def ratios(hs):                                          # smashMEM.py:58-70, 95-110
            if not hs:
                return []
            rlen = hs[0].rlen
            code = np.zeros(rlen, dtype=int)
            spans = []
            for h in hs:
                if h.rev:
                    s, e = rlen - h.qend, rlen - h.qstart
                else:
                    s, e = h.qstart, h.qend
                code[s:e] += 1
            for h in hs:
                if h.rev:
                    s, e = h.rlen - h.qend, h.rlen - h.qstart
                else:
                    s, e = h.qstart, h.qend
                spans.append(np.sum(code[s:e] == 1) / float(h.qlen))
            return spans

        ra1, ra2 = ratios(r1), ratios(r2)

        def info(h, ratio):
            return (int(h.read2) + 1, h.HI, h.chrom if h.tid != -1 else "*", h.pos, int(h.rev),
                    h.rlen, h.qstart, h.qlen, int(np.round(h.qlen * ratio)),
                    h.qlen - max(h.L0, h.R0))

        k1 = [(h, info(h, r)) for h, r in zip(r1, ra1) if r >= min_ratio]
        c1 = np.array([h.tid for h, _ in k1])
        p1 = np.array([h.pos for h, _ in k1])
        k2 = []
        for h, r in zip(r2, ra2):
            if r >= min_ratio:
                near = np.sum(np.logical_and(c1 == h.tid, np.abs(p1 - h.pos) < hit_window)) if len(k1) else 0
                if near == 0:
                    k2.append((h, info(h, r)))
        o1 = np.argsort([h.HI for h, _ in k1]) if k1 else []
        o2 = np.argsort([h.HI for h, _ in k2]) if k2 else []
        key = (tuple([k1[x][0].tid for x in o1] + [k2[x][0].tid for x in o2]),
               tuple([k1[x][0].pos for x in o1] + [k2[x][0].pos for x in o2]))
        if key not in dupes:
            dupes.add(key)
            for x in o1:
                rows.append((rid,) + k1[x][1])
            for x in o2:
                rows.append((rid,) + k2[x][1])
            n_non += 1
        else:
            n_dupe += 1
    return rows, n_dupe, n_non


_POSRE = re.compile(r"^chr(\d+|[XY]) \d+$")


def positions(rows):
    """smash_mapping.sh:29: awk '{print $4,$5}' | perl -ne 'print if /^chr(\\d+|[XY]) \\d+$/'."""
    out = []
    for r in rows:
        s = f"{r[3]} {r[4]}"
        if _POSRE.match(s):
            out.append(s)
    return out


def smash_text(rows, n_dupe, n_non):
    head = ["read_id", "read_index", "hit_index", "chrom", "pos", "reverse", "read_len",
            "hit_offset", "match_len", "umatch", "excess"]
    out = ["\t".join(head)]
    for r in rows:
        out.append("\t".join([r[0].decode()] + [str(x) for x in r[1:]]))
    out.append("%d dupes\t%d non-dupes" % (n_dupe, n_non))
    return "\n".join(out) + "\n"


def varbin(position_lines, bins_rows, chrominfo):
    """varbin.py main().  bins_rows: list of str-lists (bins.txt split on tab); chrominfo:
    {name: [name,size,offset]} (chrom_sizes.txt).  Returns (counts, total, dups, kept)."""
    starts = [int(b[2]) for b in bins_rows]
    counts = [0] * len(bins_rows)
    kept = dups = total = 0
    prev = ""
    for x in position_lines:
        a = x.rstrip().split(" ")
        c, p = a[0], a[1]
        if c.find("_") > -1 or c == "chrM" or c == "" or c not in chrominfo:
            continue
        total += 1
        if p == prev:
            dups += 1
            continue
        ab = int(p) + int(chrominfo[c][2])
        kept += 1
        counts[bisect.bisect(starts, ab) - 1] += 1
        prev = p
    return counts, total, dups, kept


def varbin_text(bins_rows, counts, kept):
    out = []
    for b, c in zip(bins_rows, counts):
        ratio = float(c) / (float(kept) / float(len(bins_rows)))
        out.append("\t".join(b[0:3]) + "\t" + str(c) + "\t" + str(ratio) + "\n")
    return "".join(out)


def varbin_stats_text(counts, total, dups, kept):
    s = sorted(counts)
    return ("TotalReads\tDupsRemoved\tReadsKept\tMedianBinCount\n"
            f"{total}\t{dups}\t{kept}\t{s[len(counts) // 2]}\n")


def read_table(path, sep="\t"):
    return [ln.rstrip().split(sep) for ln in open(path)]


def read_chrominfo(path):
    d = {}
    for r in read_table(path):
        d.setdefault(r[0], r)
    return d
