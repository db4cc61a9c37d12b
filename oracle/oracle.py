"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/liboracle.so (the plain-C restatement of the reference hot path) plus
a reader for the reference's on-disk index (`<fa>.bin/rc1.*`; written by
fasta.cpp:215-236 and longSA.cpp:179-190) and a runner for the unmodified reference binaries in
oracle/_ref.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference
legs import this module.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import shutil
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
        _LIB = C.CDLL(so)
        _LIB.orc_mam.restype = C.c_uint64
        _LIB.orc_mem.restype = C.c_uint64
        _LIB.orc_mam_bruteforce.restype = C.c_uint64
        _LIB.orc_map_batch.restype = C.c_uint64
    return _LIB


class _Index(C.Structure):
    _fields_ = [("text", C.c_void_p), ("N", C.c_uint64), ("sa", C.c_void_p), ("isa", C.c_void_p),
                ("w", C.c_int), ("lcp_vec", C.c_void_p), ("lcp_m", C.c_void_p), ("n_m", C.c_uint64),
                ("n_descr", C.c_uint64), ("startpos", C.c_void_p), ("sizes", C.c_void_p),
                ("descr", C.POINTER(C.c_char_p)), ("rcref", C.c_int)]


class _Batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("names", C.c_void_p), ("name_off", C.c_void_p),
                ("seq", C.c_void_p), ("qual", C.c_void_p), ("seq_off", C.c_void_p),
                ("opt", C.c_void_p), ("opt_off", C.c_void_p), ("read_flag", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("mode", C.c_int), ("min_len", C.c_uint32), ("nomap", C.c_int),
                ("nucleotides_only", C.c_int), ("n_threads", C.c_int)]


MUM, MAM, MEM = 0, 1, 2
MATCH_DT = np.dtype([("ref", "<u8"), ("query", "<u8"), ("len", "<u8")])
LCPM_DT = np.dtype([("idx", "<u8"), ("val", "<u8")])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class Index:
    """The reference's index, as arrays.  Built either from `<fa>.bin/` files or from a text."""

    def __init__(self, text, sa, isa, lcp_vec, lcp_m, startpos, sizes, descr, rcref=1, w=None,
                 fasta_size=0):
        self.text, self.sa, self.isa, self.lcp_vec, self.lcp_m = text, sa, isa, lcp_vec, lcp_m
        self.startpos = np.ascontiguousarray(startpos, dtype=np.uint64)
        self.sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        self.descr = list(descr)
        self.rcref = rcref
        self.N = int(len(text))
        self.w = w or sa.dtype.itemsize
        self.fasta_size = fasta_size
        self._descr_c = (C.c_char_p * len(self.descr))(*[d.encode() for d in self.descr])
        self.c = _Index(_ptr(text), self.N, _ptr(sa), _ptr(isa), self.w, _ptr(lcp_vec),
                        _ptr(lcp_m), len(lcp_m), len(self.descr), _ptr(self.startpos),
                        _ptr(self.sizes), self._descr_c, rcref)

    @property
    def logN(self):
        import math
        return int(math.ceil(math.log(self.N) / math.log(2.0)))

    # ---- on-disk format -------------------------------------------
    @staticmethod
    def load(fasta, rcref=1):
        base = f"{fasta}.bin/rc{rcref}"
        with open(base + ".ref.bin", "rb") as f:
            fasta_size, N, nd = struct.unpack("<3Q", f.read(24))
            startpos, sizes, descr = [], [], []
            for _ in range(nd):
                sp, sz, sl = struct.unpack("<3Q", f.read(24))
                startpos.append(sp); sizes.append(sz); descr.append(f.read(sl).decode())
        text = np.fromfile(base + ".ref.seq.bin", dtype=np.uint8)
        assert len(text) == N
        w = 8 if os.path.exists(base + ".i8.index.bin") and not os.path.exists(base + ".i4.index.bin") else 4
        ib = f"{base}.i{w}.index"
        hdr = np.fromfile(ib + ".bin", dtype="<u8")
        assert hdr[0] == fasta_size and hdr[3] == N
        dt = np.dtype("<u4") if w == 4 else np.dtype("<u8")
        sa = np.fromfile(ib + ".sa.bin", dtype=dt)
        isa = np.fromfile(ib + ".isa.bin", dtype=dt)
        vec = np.fromfile(ib + ".lcp.vec.bin", dtype=np.uint8)
        raw = np.fromfile(ib + ".lcp.m.bin", dtype=np.uint8)
        m = np.zeros(len(raw) // 16, dtype=LCPM_DT)
        if len(raw):
            r = raw.reshape(-1, 16)
            m["idx"] = r[:, :8].copy().view("<u8").reshape(-1)
            if w == 4:
                m["val"] = r[:, 8:12].copy().view("<u4").reshape(-1)
            else:
                m["val"] = r[:, 8:16].copy().view("<u8").reshape(-1)
        assert len(m) == hdr[5]
        return Index(text, sa, isa, vec, m, startpos, sizes, descr, rcref, w, fasta_size)

    @staticmethod
    def text_from_chromosomes(names, seqs, rcref=1):
        """The Sequence text layout of fasta.cpp:151-203: fwd ` rc ` ... $ (all lower case)."""
        comp = np.arange(256, dtype=np.uint8)
        for a, b in zip(b"acgtrymkbdhv", b"tgcayrkmvhdb"):
            comp[a] = b
        parts, startpos, sizes, descr = [], [], [], []
        pos = 0
        for k, (n, s) in enumerate(zip(names, seqs)):
            low = np.where((s >= 65) & (s <= 90), s + 32, s).astype(np.uint8)
            last = k == len(names) - 1
            startpos.append(pos); sizes.append(len(low)); descr.append(n)
            parts.append(low); pos += len(low)
            if rcref or not last:
                parts.append(np.array([0x60], dtype=np.uint8)); pos += 1
            if rcref:
                startpos.append(pos); sizes.append(len(low)); descr.append(n)
                parts.append(comp[low[::-1]]); pos += len(low)
                if not last:
                    parts.append(np.array([0x60], dtype=np.uint8)); pos += 1
        parts.append(np.array([0x24], dtype=np.uint8))
        return np.concatenate(parts), startpos, sizes, descr

    @staticmethod
    def build(names, seqs, rcref=1, w=4):
        """Small-text index through the oracle's own comparison sort (tests only)."""
        text, startpos, sizes, descr = Index.text_from_chromosomes(names, seqs, rcref)
        N = len(text)
        sa = np.zeros(N, dtype=np.uint64); isa = np.zeros(N, dtype=np.uint64)
        lcp = np.zeros(N, dtype=np.uint64)
        lib().orc_build_index(_ptr(text), C.c_uint64(N), _ptr(sa), _ptr(isa), _ptr(lcp))
        vec = np.minimum(lcp, 255).astype(np.uint8)
        big = np.nonzero(lcp >= 255)[0]
        m = np.zeros(len(big), dtype=LCPM_DT)
        m["idx"] = big; m["val"] = lcp[big]
        dt = np.uint32 if w == 4 else np.uint64
        return Index(text, sa.astype(dt), isa.astype(dt), vec, m, startpos, sizes, descr, rcref, w)

    def save(self, fasta):
        """Write the reference's file set for this index (format: SURVEY.md Appendix B)."""
        self.fasta_size = os.path.getsize(fasta)
        os.makedirs(fasta + ".bin", exist_ok=True)
        base = f"{fasta}.bin/rc{self.rcref}"
        with open(base + ".ref.bin", "wb") as f:
            f.write(struct.pack("<3Q", self.fasta_size, self.N, len(self.descr)))
            for sp, sz, d in zip(self.startpos, self.sizes, self.descr):
                f.write(struct.pack("<3Q", int(sp), int(sz), len(d))); f.write(d.encode())
            f.write(struct.pack("<Q", max(len(d) for d in self.descr)))
        self.text.tofile(base + ".ref.seq.bin")
        ib = f"{base}.i{self.w}.index"
        np.array([self.fasta_size, self.logN, self.N - 1, self.N, self.N, len(self.lcp_m)],
                 dtype="<u8").tofile(ib + ".bin")
        self.sa.tofile(ib + ".sa.bin"); self.isa.tofile(ib + ".isa.bin")
        self.lcp_vec.tofile(ib + ".lcp.vec.bin")
        raw = np.zeros((len(self.lcp_m), 16), dtype=np.uint8)
        if len(self.lcp_m):
            raw[:, :8] = self.lcp_m["idx"].astype("<u8").view(np.uint8).reshape(-1, 8)
            raw[:, 8:16] = self.lcp_m["val"].astype("<u8").view(np.uint8).reshape(-1, 8)
            if self.w == 4:
                raw[:, 12:] = 0
        raw.tofile(ib + ".lcp.m.bin")

    # ---- searches ---------------------------------------------------------------------------
    def _search(self, fn, query: bytes, min_len):
        q = np.frombuffer(query.lower(), dtype=np.uint8)
        out = np.zeros(max(16, 4 * len(q)), dtype=MATCH_DT)
        n = fn(C.byref(self.c), _ptr(q), C.c_uint64(len(q)), C.c_uint64(min_len), _ptr(out),
               C.c_uint64(len(out)))
        if n > len(out):
            out = np.zeros(n, dtype=MATCH_DT)
            n = fn(C.byref(self.c), _ptr(q), C.c_uint64(len(q)), C.c_uint64(min_len), _ptr(out),
                   C.c_uint64(len(out)))
        return out[:n]

    def mam(self, query, min_len=20):
        return self._search(lib().orc_mam, query, min_len)

    def mem(self, query, min_len=20):
        return self._search(lib().orc_mem, query, min_len)

    def mam_bruteforce(self, query, min_len=20):
        q = np.frombuffer(query.lower(), dtype=np.uint8)
        out = np.zeros(len(q) + 1, dtype=MATCH_DT)
        n = lib().orc_mam_bruteforce(_ptr(self.text), C.c_uint64(self.N), _ptr(q),
                                     C.c_uint64(len(q)), C.c_uint64(min_len), _ptr(out),
                                     C.c_uint64(len(out)))
        return out[:n]

    def mappability(self):
        """map.bin body (without the 2 junk header bytes)."""
        out = np.zeros(2 * int(self.sizes[::2].sum()), dtype=np.uint8)
        lib().orc_mappability(C.byref(self.c), _ptr(out))
        return out

    def sam_header(self):
        """Sequence::sam_header (fasta.cpp:243-252)."""
        step = 2 if self.rcref else 1
        s = "@HD\tVN:1.0\tSO:unsorted\n"
        for i in range(0, len(self.descr), step):
            s += f"@SQ\tSN:{self.descr[i]}\tLN:{int(self.sizes[i])}\n"
        return s + "@PG\tID:longMEM\tPN:longMEM\tVN:0.5\n"

    def map_batch(self, batch, mode=MAM, min_len=20, nomap=True, nucleotides_only=False,
                  n_threads=1, want_matches=False):
        """mummer -samin -samout over a packed batch -> SAM record bytes (input order)."""
        rf = read_flags(batch)
        b = _Batch(batch.n, _ptr(batch.names), _ptr(batch.name_off), _ptr(batch.seq),
                   _ptr(batch.qual), _ptr(batch.seq_off), _ptr(batch.opt), _ptr(batch.opt_off),
                   _ptr(rf))
        if batch.opt.size == 0:
            b.opt = None
        p = _Params(mode, min_len, int(nomap), int(nucleotides_only), n_threads)
        cap = int(batch.n) * 2600 + 4096
        moff = np.zeros(batch.n + 1, dtype=np.int64) if want_matches else None
        mcap = int(batch.n) * 24 + 64
        mbuf = np.zeros(mcap, dtype=MATCH_DT) if want_matches else None
        while True:
            out = np.empty(cap, dtype=np.uint8)
            need = lib().orc_map_batch(C.byref(self.c), C.byref(b), C.byref(p), _ptr(out),
                                       C.c_uint64(cap), _ptr(moff) if want_matches else None,
                                       _ptr(mbuf) if want_matches else None, C.c_uint64(mcap))
            ok = need <= cap and (not want_matches or moff[-1] <= mcap)
            if ok:
                break
            cap = max(cap, int(need) + 16)
            if want_matches and moff[-1] > mcap:
                mcap = int(moff[-1]) + 16
                mbuf = np.zeros(mcap, dtype=MATCH_DT)
        sam = out[:need].tobytes()
        if want_matches:
            return sam, moff, mbuf[:moff[-1]]
        return sam


def read_flags(batch):
    """QueryReader::run + Aligner::reset (query.cpp:643-644, 185-201): name gets ':0'/':1' from
    flag 64/128, then a trailing ':0'/':1' is stripped into read_flag 65/129.  batch.names never
    carry the suffix, so only names that *themselves* end in :0/:1 need the strip quirk."""
    pre = getattr(batch, "read_flag", None)
    if pre is not None:
        return np.ascontiguousarray(pre, dtype=np.uint16)
    fl = batch.flags.astype(np.uint16)
    rf = np.where(fl & 64, 65, np.where(fl & 128, 129, 0)).astype(np.uint16)
    return rf


# ---- the unmodified reference (oracle/_ref) -------------------------------------------------

def have_reference():
    return os.path.exists(os.path.join(REF_BIN, "mummer"))


def ref_build_index(fasta, long_ints=False, mappability=True, verbose=False):
    """index_setup.sh:19,22 with the compiled reference; exits 1 by design on 'dummy'."""
    exe = os.path.join(REF_BIN, "mummer-long" if long_ints else "mummer")
    wd = os.path.dirname(os.path.abspath(fasta))
    subprocess.run([exe, "-rcref", fasta, "dummy"], cwd=wd, stdout=subprocess.DEVNULL,
                   stderr=None if verbose else subprocess.DEVNULL)
    if mappability:
        subprocess.run([exe, "-rcref", "-mappability", fasta, fasta + ".bin/map.bin"], cwd=wd,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def ref_map(fasta, reads_sam, workdir, threads=2, extra=(), long_ints=False):
    """smash_mapping.sh:19 -> (header bytes, sorted list of record lines)."""
    exe = os.path.join(REF_BIN, "mummer-long" if long_ints else "mummer")
    shutil.rmtree(os.path.join(workdir, "mapout"), ignore_errors=True)
    subprocess.run([exe, "-rcref", "-qthreads", str(max(2, threads)), "-nomap", "-samin", "-samout",
                    *extra, fasta, reads_sam], cwd=workdir, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    header, lines = None, []
    for fn in sorted(glob.glob(os.path.join(workdir, "mapout", "*.txt"))):
        h = []
        with open(fn, "rb") as f:
            for ln in f:
                (h if ln.startswith(b"@") else lines).append(ln)
        hb = b"".join(h)
        assert header is None or header == hb
        header = hb
    return header, sorted(lines)


def ref_map_chunks(fasta, reads_sam, workdir, threads=2, extra=(), long_ints=False):
    """Like ref_map, but every chunk file as the reference wrote it: list of record-line lists (file order kept)."""
    exe = os.path.join(REF_BIN, "mummer-long" if long_ints else "mummer")
    shutil.rmtree(os.path.join(workdir, "mapout"), ignore_errors=True)
    subprocess.run([exe, "-rcref", "-qthreads", str(max(2, threads)), "-nomap", "-samin", "-samout", *extra, fasta, reads_sam],
                   cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = []
    for fn in sorted(glob.glob(os.path.join(workdir, "mapout", "*.txt"))):
        with open(fn, "rb") as f:
            out.append([ln for ln in f if not ln.startswith(b"@")])
    return out


def memsam_sort_key(names, sizes):
    """Key function for SAM record lines reproducing MemSam::operator< (memsam.h:136-158) with MemSam::chromosomes
    as Pairs::Pairs fills it (query.cpp:546-552): (offset[RNAME] + POS, name bytes, flag & (64|128|16))."""
    off, acc = {}, 0
    for n, s in zip(names, sizes):
        off[n.encode() if isinstance(n, str) else n] = acc
        acc += int(s)
    off[b"*"] = acc

    def key(line):
        f = line.split(b"\t", 4)
        return off[f[2]] + int(f[3]), f[0], int(f[1]) & (64 | 128 | 16)
    return key


def ref_mappability_tag(fasta, sam_path):
    """smash_mapping.sh:23 first stage; returns stdout bytes."""
    exe = os.path.join(REF_BIN, "mappability_tag")
    return subprocess.run([exe, fasta, sam_path], check=True, stdout=subprocess.PIPE).stdout
