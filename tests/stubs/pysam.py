"""tests/stubs/pysam.py -- TEST INFRASTRUCTURE ONLY: a stand-in for the `pysam` module, just large enough to run the
UNMODIFIED /root/reference/smashMEM.py under python3 (pysam and samtools are not installable in this image).

It is put on PYTHONPATH only by tests/golden/make_golden_smash.py, which runs the reference script as a subprocess
and commits its stdout as golden vectors.  Nothing else imports it.

What it provides is exactly what smashMEM.py touches (smashMEM.py:12-52, 58-110, 181-202):

  Samfile(path, mode)      records of a SAM *text* file in file order (the caller name-sorts it the way
                           `samtools sort -n` does, smash_mapping.sh:23), next(), reset(), getrname(tid), close()
  AlignedRead              qname, is_read1, is_read2, is_unmapped, is_reverse, tid, pos, rlen, qstart, qend, qlen, opt(tag)

Accessor semantics follow the pysam documentation of the AlignedRead API the script was written against (pysam 0.7/0.8,
csamtools.pyx): pos = 0-based leftmost coordinate (POS-1); tid = index of RNAME among the @SQ lines, -1 for '*';
rlen = length of the read sequence (l_qseq); qstart = start index of the aligned query portion (soft-clipped bases at
the start are skipped, hard clips ignored); qend = end index of the aligned query portion (l_qseq minus the soft clip
at the end); qlen = qend - qstart; opt(tag) = value of the optional field, KeyError when absent.
"""
import re

_CIG = re.compile(r"(\d+)([MIDNSHP=X])")


class AlignedRead(object):
    __slots__ = ("qname", "flag", "tid", "pos", "rlen", "qstart", "qend", "qlen", "_tags")

    def __init__(self, line, tid_of):
        f = line.rstrip("\n").split("\t")
        self.qname = f[0]
        self.flag = int(f[1])
        self.tid = tid_of.get(f[2], -1)
        self.pos = int(f[3]) - 1
        self.rlen = 0 if f[9] == "*" else len(f[9])
        ops = [(int(n), op) for n, op in _CIG.findall(f[5])] if f[5] != "*" else []
        start = 0
        for n, op in ops:
            if op == "H":
                continue
            if op == "S":
                start += n
            else:
                break
        end = self.rlen
        for n, op in reversed(ops):
            if op == "H":
                continue
            if op == "S":
                end -= n
            else:
                break
        self.qstart, self.qend = start, end
        self.qlen = end - start
        self._tags = {}
        for t in f[11:]:
            tag, typ, val = t.split(":", 2)
            self._tags[tag] = int(val) if typ == "i" else float(val) if typ == "f" else val

    is_paired = property(lambda s: bool(s.flag & 1))
    is_unmapped = property(lambda s: bool(s.flag & 4))
    mate_is_unmapped = property(lambda s: bool(s.flag & 8))
    is_reverse = property(lambda s: bool(s.flag & 16))
    is_read1 = property(lambda s: bool(s.flag & 64))
    is_read2 = property(lambda s: bool(s.flag & 128))
    is_secondary = property(lambda s: bool(s.flag & 256))

    def opt(self, tag):
        return self._tags[tag]


class Samfile(object):
    def __init__(self, path, mode="r"):
        self._path = path
        self._names = []
        with open(path) as f:
            for line in f:
                if not line.startswith("@"):
                    break
                if line.startswith("@SQ"):
                    for t in line.rstrip("\n").split("\t")[1:]:
                        if t.startswith("SN:"):
                            self._names.append(t[3:])
        self._tid_of = {n: i for i, n in enumerate(self._names)}
        self._f = None
        self.reset()

    def reset(self):
        if self._f:
            self._f.close()
        self._f = open(self._path)

    def __iter__(self):
        return self

    def __next__(self):
        for line in self._f:
            if line.startswith("@"):
                continue
            return AlignedRead(line, self._tid_of)
        raise StopIteration

    next = __next__

    def getrname(self, tid):
        return self._names[tid]

    def close(self):
        if self._f:
            self._f.close()
            self._f = None
