"""ctypes binding of libsmash_b200.so (include/smash_b200.h) -- the product path.

No CPU fallback: if the library is missing, or no sm_100 GPU is present, constructing a
`Context` raises.  The oracle is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMASH_B200_LIB") or os.path.join(_HERE, "libsmash_b200.so")   # override: kernel A/B builds
MODE_MUM, MODE_MAM, MODE_MEM = 0, 1, 2
WANT_SAM, WANT_MATCHES, WANT_TAIL, WANT_SORTED = 1, 2, 4, 8
N_SLOTS = 4


class SmashError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("device", C.c_int), ("mode", C.c_int), ("min_len", C.c_uint32), ("nomap", C.c_int),
                ("nucleotides_only", C.c_int), ("tag_mappability", C.c_int),
                ("max_batch_reads", C.c_uint64), ("seed_k", C.c_int)]


class _Batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("names", C.c_void_p), ("name_off", C.c_void_p),
                ("seq", C.c_void_p), ("qual", C.c_void_p), ("seq_off", C.c_void_p),
                ("opt", C.c_void_p), ("opt_off", C.c_void_p), ("read_flag", C.c_void_p),
                ("first_pair_ordinal", C.c_uint64)]


class _Result(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_matches", C.c_uint64), ("n_records", C.c_uint64),
                ("sam_bytes", C.c_uint64), ("sam", C.c_void_p), ("match_off", C.c_void_p),
                ("matches", C.c_void_p), ("gpu_ms", C.c_float)]


class _TailStats(C.Structure):
    _fields_ = [("total_reads", C.c_uint64), ("dups_removed", C.c_uint64), ("reads_kept", C.c_uint64),
                ("n_dupe_pairs", C.c_uint64), ("n_non_dupe_pairs", C.c_uint64), ("n_positions", C.c_uint64)]


class _Text(C.Structure):
    _fields_ = [("kind", C.c_int), ("flags", C.c_int), ("text", C.c_void_p * 2), ("n_bytes", C.c_uint64 * 2),
                ("first_pair_ordinal", C.c_uint64)]


class _TextInfo(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("consumed", C.c_uint64 * 2), ("mate2_first_next", C.c_int)]


TEXT_SAM, TEXT_FASTQ_PAIR = 0, 1
TEXT_FINAL, TEXT_REPLACE_N, TEXT_MATE2_FIRST = 1, 2, 4


class DeviceBatch:
    """Host copy of the packed batch a slot holds in HBM (smash_fetch_batch)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def n(self):
        return len(self.read_flag)


class _TailEdge(C.Structure):
    _fields_ = [("n_filtered", C.c_uint64), ("first_pos", C.c_int64), ("last_pos", C.c_int64)]


_lib = None


def load_library():
    """Load libsmash_b200.so; fails loudly when it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmashError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.smash_last_error.restype = C.c_char_p
    L.smash_index_text_len.restype = C.c_uint64
    L.smash_index_sam_header.restype = C.c_size_t
    L.smash_ctx_launch_count.restype = C.c_uint64
    L.smash_ctx_index_bytes.restype = C.c_uint64
    L.smash_ctx_stream.restype = C.c_void_p
    L.smash_ctx_ingest_ms.restype = C.c_double
    L.smash_host_alloc.restype = C.c_void_p
    L.smash_host_alloc.argtypes = [C.c_size_t]
    L.smash_host_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise SmashError(f"libsmash_b200 error {rc}: {load_library().smash_last_error().decode(errors='replace')}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def memcpy(dst_ptr, src_ptr, nbytes):
    """cudaMemcpyDefault between any two pointers (host or device)."""
    _check(load_library().smash_memcpy(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), C.c_size_t(nbytes)))


def device_count():
    return int(load_library().smash_device_count())


def read_flags_from_sam_flags(flags, names=None, name_off=None):
    """QueryReader::run + Aligner::reset (query.cpp:643-644, 185-201) -> 0 / 65 / 129."""
    fl = np.asarray(flags).astype(np.uint16)
    return np.where(fl & 64, 65, np.where(fl & 128, 129, 0)).astype(np.uint16)


def _read_flag(batch):
    rf = getattr(batch, "read_flag", None)
    return np.ascontiguousarray(rf, dtype=np.uint16) if rf is not None else read_flags_from_sam_flags(batch.flags)


class PinnedArray:
    """numpy view over cudaHostAlloc memory (so H2D copies of batches are true async DMA)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * self.dtype.itemsize
        self._p = load_library().smash_host_alloc(max(n, 16))
        if not self._p:
            raise SmashError("cudaHostAlloc failed")
        buf = (C.c_uint8 * max(n, 16)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            load_library().smash_host_free(C.c_void_p(self._p))
            self._p = None


class Index:
    """The reference's `<fa>.bin/` index (Sequence + longSA load branches)."""

    def __init__(self, handle, keep=None):
        self.h = handle
        self._keep = keep

    @staticmethod
    def open(fasta, rcref=True):
        L = load_library()
        h = C.c_void_p()
        _check(L.smash_index_open(str(fasta).encode(), int(bool(rcref)), C.byref(h)))
        return Index(h)

    @staticmethod
    def from_arrays(text, sa, isa, lcp_vec, lcp_m_raw, startpos, sizes, descr, rcref=True):
        L = load_library()
        h = C.c_void_p()
        startpos = np.ascontiguousarray(startpos, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        d = (C.c_char_p * len(descr))(*[x.encode() for x in descr])
        n_m = 0 if lcp_m_raw is None else lcp_m_raw.size // 16
        _check(L.smash_index_from_arrays(_ptr(text), C.c_uint64(len(text)), _ptr(sa), _ptr(isa) if isa is not None else None,
                                         int(sa.dtype.itemsize), _ptr(lcp_vec), _ptr(lcp_m_raw) if n_m else None,
                                         C.c_uint64(n_m), C.c_uint64(len(descr)), _ptr(startpos), _ptr(sizes), d,
                                         int(bool(rcref)), C.byref(h)))
        return Index(h, keep=(text, sa, isa, lcp_vec, lcp_m_raw, startpos, sizes, d))

    @property
    def text_len(self):
        return int(load_library().smash_index_text_len(self.h))

    @property
    def int_width(self):
        """4 or 8: which of rc1.i4.index.* / rc1.i8.index.* was opened (size.h:9-22)."""
        return int(load_library().smash_index_int_width(self.h))

    def sam_header(self):
        L = load_library()
        n = L.smash_index_sam_header(self.h, None, C.c_size_t(0))
        buf = C.create_string_buffer(n)
        L.smash_index_sam_header(self.h, buf, C.c_size_t(n))
        return buf.raw[:n]

    def close(self):
        if self.h:
            load_library().smash_index_close(self.h)
            self.h = None


class Result:
    def __init__(self, r: _Result, want):
        self.n_reads = int(r.n_reads)
        self.sam_bytes = int(r.sam_bytes)
        self.gpu_ms = float(r.gpu_ms)
        self.sam = None
        self.match_off = None
        self.matches = None
        if (want & WANT_SAM) and r.sam:
            self.sam = C.string_at(r.sam, self.sam_bytes)
        if (want & WANT_MATCHES) and r.match_off:
            n = self.n_reads
            self.match_off = np.ctypeslib.as_array(C.cast(r.match_off, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
            m = int(self.match_off[-1])
            self.matches = (np.ctypeslib.as_array(C.cast(r.matches, C.POINTER(C.c_uint64)), shape=(max(m, 1) * 3,))
                            [:3 * m].reshape(-1, 3).copy())


class Context:
    """One GPU: index in HBM + batch slots (Pairs/Pair/Aligner of the reference)."""

    def __init__(self, index: Index, device=0, mode=MODE_MAM, min_len=20, nomap=True,
                 nucleotides_only=False, tag_mappability=False, seed_k=0):
        L = load_library()
        if L.smash_device_count() <= 0:
            raise SmashError("no sm_100 CUDA device: libsmash_b200 has no CPU fallback")
        p = _Params()
        L.smash_params_default(C.byref(p))
        p.device, p.mode, p.min_len, p.nomap = device, mode, min_len, int(nomap)
        p.nucleotides_only, p.tag_mappability, p.seed_k = int(nucleotides_only), int(tag_mappability), seed_k
        self.h = C.c_void_p()
        self.index = index
        _check(L.smash_ctx_create(index.h, C.byref(p), C.byref(self.h)))
        self._inflight = {}

    @classmethod
    def from_text(cls, text, startpos, sizes, descr, rcref=True, w=None, keep_isa=True, chunk_cap=0,
                  device=0, mode=MODE_MAM, min_len=20, nomap=True, nucleotides_only=False,
                  tag_mappability=False, seed_k=0):
        """Build the index ON THE GPU from the Sequence text (longSA build branch) and keep it in HBM."""
        L = load_library()
        if L.smash_device_count() <= 0:
            raise SmashError("no sm_100 CUDA device: libsmash_b200 has no CPU fallback")
        p = _Params()
        L.smash_params_default(C.byref(p))
        p.device, p.mode, p.min_len, p.nomap = device, mode, min_len, int(nomap)
        p.nucleotides_only, p.tag_mappability, p.seed_k = int(nucleotides_only), int(tag_mappability), seed_k
        text = np.ascontiguousarray(text, dtype=np.uint8)
        sp = np.ascontiguousarray(startpos, dtype=np.uint64)
        sz = np.ascontiguousarray(sizes, dtype=np.uint64)
        d = (C.c_char_p * len(descr))(*[x.encode() for x in descr])
        if w is None:
            w = 4 if len(text) < 0xFFFFFFFF - 100000 else 8          # mummer.cpp:156-183
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        self.index = None
        self._inflight = {}
        self.N, self.w, self.sizes = len(text), w, sz
        _check(L.smash_ctx_create_from_text(_ptr(text), C.c_uint64(len(text)), C.c_uint64(len(descr)), _ptr(sp),
                                            _ptr(sz), d, int(bool(rcref)), int(w), int(bool(keep_isa)),
                                            C.c_uint64(chunk_cap), C.byref(p), C.byref(self.h)))
        return self

    def copy_index(self, N, w, want_isa=True):
        dt = np.uint32 if w == 4 else np.uint64
        sa = np.empty(N, dtype=dt)
        isa = np.empty(N, dtype=dt) if want_isa else None
        vec = np.empty(N, dtype=np.uint8)
        nm = C.c_uint64()
        _check(load_library().smash_ctx_copy_index(self.h, _ptr(sa), _ptr(isa) if want_isa else None, _ptr(vec), None, C.byref(nm)))
        m = np.zeros(16 * max(nm.value, 1), dtype=np.uint8)
        _check(load_library().smash_ctx_copy_index(self.h, None, None, None, _ptr(m), None))
        return sa, isa, vec, m[:16 * nm.value]

    def build_mappability_device(self, total_forward_bases=None):
        """map.bin body built and kept in HBM only (no host copy)."""
        _check(load_library().smash_ctx_build_mappability(self.h, None, C.c_uint64(0)))

    def drop_isa(self):
        _check(load_library().smash_ctx_drop_isa(self.h))

    def save_index(self, fasta, with_mappability=False):
        _check(load_library().smash_ctx_save_index(self.h, str(fasta).encode(), int(with_mappability)))

    def close(self):
        if self.h:
            load_library().smash_ctx_destroy(self.h)
            self.h = None

    # -- mappability -----------------------------------------------------------------------
    def load_mappability(self, body: np.ndarray):
        body = np.ascontiguousarray(body, dtype=np.uint8)
        _check(load_library().smash_ctx_load_mappability(self.h, _ptr(body), C.c_uint64(body.size)))

    def load_mappability_file(self, path):
        self.load_mappability(np.fromfile(path, dtype=np.uint8)[2:])     # 2 junk bytes (longSA.cpp:606-617)

    def build_mappability(self, total_forward_bases):
        out = np.empty(2 * int(total_forward_bases), dtype=np.uint8)
        _check(load_library().smash_ctx_build_mappability(self.h, _ptr(out), C.c_uint64(out.size)))
        return out

    # -- batches ---------------------------------------------------------------------------
    @staticmethod
    def _cbatch(batch, read_flag, first_pair=0):
        b = _Batch(batch.n, _ptr(batch.names), _ptr(batch.name_off), _ptr(batch.seq), _ptr(batch.qual),
                   _ptr(batch.seq_off), _ptr(batch.opt) if batch.opt.size else None,
                   _ptr(batch.opt_off) if batch.opt.size else None, _ptr(read_flag), first_pair)
        return b

    def map_batch(self, batch, want=WANT_SAM, first_pair=0, read_flag=None):
        rf = _read_flag(batch) if read_flag is None else read_flag
        b = self._cbatch(batch, rf, first_pair)
        r = _Result()
        _check(load_library().smash_map_batch(self.h, C.byref(b), want, C.byref(r)))
        return Result(r, want)

    def submit(self, slot, batch, want=WANT_SAM, first_pair=0, read_flag=None):
        rf = _read_flag(batch) if read_flag is None else read_flag
        b = self._cbatch(batch, rf, first_pair)
        self._inflight[slot] = (batch, rf, b, want)
        _check(load_library().smash_submit(self.h, slot, C.byref(b), want))

    def wait(self, slot, copy=True):
        r = _Result()
        _check(load_library().smash_wait(self.h, slot, C.byref(r)))
        want = self._inflight.pop(slot)[3]
        if not copy:
            return r
        return Result(r, want)

    def upload(self, batch, read_flag=None):
        rf = _read_flag(batch) if read_flag is None else read_flag
        b = self._cbatch(batch, rf)
        self._resident = (batch, rf)
        _check(load_library().smash_batch_upload(self.h, C.byref(b)))

    def map_resident(self, want=WANT_SAM):
        r = _Result()
        _check(load_library().smash_map_resident(self.h, want, C.byref(r)))
        return r

    # -- input side on the device (raw SAM text / FASTQ pair text in; ingest.cu) ------------------------
    @staticmethod
    def _ctext(kind, text0, text1, final, replace_n, first_pair, mate2_first=False):
        """text: bytes or a uint8 numpy array (e.g. PinnedArray.array); returns (struct, objects to keep alive)."""
        t = _Text()
        t.kind = kind
        t.flags = (TEXT_FINAL if final else 0) | (TEXT_REPLACE_N if replace_n else 0) | (TEXT_MATE2_FIRST if mate2_first else 0)
        keep = []
        for f, x in enumerate((text0, text1 if kind == TEXT_FASTQ_PAIR else b"")):
            if isinstance(x, np.ndarray):
                x = np.ascontiguousarray(x, dtype=np.uint8)
                t.text[f], t.n_bytes[f] = (x.ctypes.data if x.size else None), x.size
            else:
                x = bytes(x)
                t.text[f], t.n_bytes[f] = (C.cast(C.c_char_p(x), C.c_void_p).value if len(x) else None), len(x)
            keep.append(x)
        t.first_pair_ordinal = first_pair
        return t, keep

    def text_upload(self, kind, text0, text1=b"", final=True, replace_n=False, first_pair=0, mate2_first=False):
        """Parse raw text on the GPU into slot 0's resident batch -> (n_reads, (consumed0, consumed1)).
        self.mate2_first_next: pass it as mate2_first with the next chunk of a FASTQ pair stream."""
        t, keep = self._ctext(kind, text0, text1, final, replace_n, first_pair, mate2_first)
        info = _TextInfo()
        _check(load_library().smash_text_upload(self.h, C.byref(t), C.byref(info)))
        self.mate2_first_next = bool(info.mate2_first_next)
        return int(info.n_reads), (int(info.consumed[0]), int(info.consumed[1]))

    def submit_text(self, slot, kind, text0, text1=b"", final=True, replace_n=False, first_pair=0, want=WANT_SAM, mate2_first=False):
        t, keep = self._ctext(kind, text0, text1, final, replace_n, first_pair, mate2_first)
        info = _TextInfo()
        _check(load_library().smash_submit_text(self.h, slot, C.byref(t), want, C.byref(info)))
        self.mate2_first_next = bool(info.mate2_first_next)
        self._inflight[slot] = (keep, None, t, want)
        return int(info.n_reads), (int(info.consumed[0]), int(info.consumed[1]))

    def ingest_ms(self, reset=False):
        return float(load_library().smash_ctx_ingest_ms(self.h, int(reset)))

    def fetch_batch(self, slot=0):
        L = load_library()
        n, nb, sb, ob = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(L.smash_batch_sizes(self.h, slot, C.byref(n), C.byref(nb), C.byref(sb), C.byref(ob)))
        n, nb, sb, ob = n.value, nb.value, sb.value, ob.value
        names = np.zeros(nb, np.uint8); seq = np.zeros(sb, np.uint8); qual = np.zeros(sb, np.uint8); opt = np.zeros(ob, np.uint8)
        name_off = np.zeros(n + 1, np.int64); seq_off = np.zeros(n + 1, np.int64); opt_off = np.zeros(n + 1, np.int64)
        rf = np.zeros(n, np.uint16)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)          # noqa: E731 -- empty arrays still need a valid pointer
        _check(L.smash_fetch_batch(self.h, slot, vp(names), vp(name_off), vp(seq), vp(qual), vp(seq_off), vp(opt), vp(opt_off), vp(rf)))
        return DeviceBatch(names=names, name_off=name_off, seq=seq, qual=qual, seq_off=seq_off, opt=opt, opt_off=opt_off,
                           read_flag=rf, flags=np.zeros(n, np.uint16))

    def fetch_sam(self):
        p = C.c_void_p()
        n = C.c_uint64()
        _check(load_library().smash_fetch_sam(self.h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    # -- tail ------------------------------------------------------------------------------
    def tail_configure(self, bin_starts, chrom_names, chrom_offsets, hit_window=10000, min_excess=4):
        bs = np.ascontiguousarray(bin_starts, dtype=np.int64)
        co = np.ascontiguousarray(chrom_offsets, dtype=np.int64)
        names = (C.c_char_p * len(chrom_names))(*[c.encode() for c in chrom_names])
        self.n_bins = len(bs)
        _check(load_library().smash_tail_configure(self.h, _ptr(bs), C.c_uint64(len(bs)), names, _ptr(co),
                                                   C.c_uint64(len(co)), C.c_int64(hit_window), C.c_int32(min_excess)))

    def tail_finish(self, counts_device_ptr=None):
        counts = np.zeros(self.n_bins, dtype=np.int64)
        st = _TailStats()
        _check(load_library().smash_tail_finish(self.h, _ptr(counts), C.c_void_p(counts_device_ptr) if counts_device_ptr else None,
                                                C.byref(st)))
        return counts, {k: int(getattr(st, k)) for k, _ in _TailStats._fields_}

    def tail_positions(self):
        c, p, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _check(load_library().smash_tail_positions(self.h, C.byref(c), C.byref(p), C.byref(n)))
        m = n.value
        if not m:
            return np.zeros(0, np.int32), np.zeros(0, np.int64)
        return (np.ctypeslib.as_array(C.cast(c, C.POINTER(C.c_int32)), shape=(m,)).copy(),
                np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), shape=(m,)).copy())

    # -- read-sharded multi-GPU tail (see multigpu.py) -----------------------------------------
    def tail_export_keys(self, ordinal_base):
        """-> (device pointer, n): {fp1, fp2, ordinal} u64 triples of this rank's dupe-set pairs."""
        ptr, n = C.c_void_p(), C.c_uint64()
        _check(load_library().smash_tail_export_keys(self.h, C.c_uint64(ordinal_base), C.byref(ptr), C.byref(n)))
        return ptr.value, int(n.value)

    def tail_phase_a(self, ordinal_base, foreign_ptr=None, n_foreign=0):
        e = _TailEdge()
        _check(load_library().smash_tail_phase_a(self.h, C.c_uint64(ordinal_base), C.c_void_p(foreign_ptr) if n_foreign else None,
                                                 C.c_uint64(n_foreign), C.byref(e)))
        return int(e.n_filtered), int(e.first_pos), int(e.last_pos)

    def tail_phase_a_verdict(self, ordinal_base, min_ord_ptr, n_keys):
        e = _TailEdge()
        _check(load_library().smash_tail_phase_a_verdict(self.h, C.c_uint64(ordinal_base), C.c_void_p(min_ord_ptr) if n_keys else None,
                                                         C.c_uint64(n_keys), C.byref(e)))
        return int(e.n_filtered), int(e.first_pos), int(e.last_pos)

    def tail_phase_b(self, has_prev=False, prev_last_pos=0, counts_device_ptr=None):
        counts = np.zeros(self.n_bins, dtype=np.int64)
        st = _TailStats()
        _check(load_library().smash_tail_phase_b(self.h, int(bool(has_prev)), C.c_int64(prev_last_pos), _ptr(counts),
                                                 C.c_void_p(counts_device_ptr) if counts_device_ptr else None, C.byref(st)))
        return counts, {k: int(getattr(st, k)) for k, _ in _TailStats._fields_}

    # -- NCCL behind the ABI (csrc/comm.cu) -------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        _check(load_library().smash_comm_unique_id(buf, C.c_size_t(128)))
        return buf.raw

    def comm_init_rank(self, rank, world, unique_id: bytes):
        _check(load_library().smash_comm_init_rank(self.h, int(rank), int(world), C.c_char_p(unique_id)))

    def bins_finish(self, ordinal_base=0, counts_device_ptr=None):
        """Collective: global bin counts + stats on every rank (partitioned dedupe exchange + ONE ncclAllReduce)."""
        counts = np.zeros(self.n_bins, dtype=np.int64)
        st = _TailStats()
        _check(load_library().smash_bins_finish(self.h, C.c_uint64(ordinal_base), _ptr(counts),
                                                C.c_void_p(counts_device_ptr) if counts_device_ptr else None, C.byref(st)))
        return counts, {k: int(getattr(st, k)) for k, _ in _TailStats._fields_}

    def tail_reserve(self, max_pairs, max_hits):
        _check(load_library().smash_tail_reserve(self.h, C.c_uint64(max_pairs), C.c_uint64(max_hits)))

    def tail_reset(self):
        _check(load_library().smash_tail_reset(self.h))

    def set_tag_mappability(self, on):
        _check(load_library().smash_ctx_set_tag_mappability(self.h, int(bool(on))))

    def set_chunking(self, max_chunks=4, min_reads=65536):
        """smash_ctx_set_chunking: how smash_submit pipelines one batch (output is identical either way)."""
        _check(load_library().smash_ctx_set_chunking(self.h, int(max_chunks), C.c_uint64(int(min_reads))))

    def set_transport(self, full_sam_text=False, host_threads=0, compact_only=False, mode=None):
        """smash_ctx_set_transport: 0 = per read range whichever way is faster (default), 1 = whole SAM text over PCIe,
        2 = compact transport + host line building only."""
        if mode is None:
            mode = 2 if compact_only else int(bool(full_sam_text))
        _check(load_library().smash_ctx_set_transport(self.h, int(mode), int(host_threads)))     # mode 3 (tests): ranges alternate

    def io_bytes(self, reset=False):
        """(h2d, d2h) bytes the library copied since the last reset."""
        a, b = C.c_uint64(), C.c_uint64()
        load_library().smash_ctx_io_bytes(self.h, C.byref(a), C.byref(b), int(reset))
        return int(a.value), int(b.value)

    def stage_ms(self, reset=False):
        out = (C.c_double * 8)()
        load_library().smash_ctx_stage_ms(self.h, out, int(reset))
        return dict(zip(["search", "records", "sizes_scan", "emit_text", "match_csr", "tail", "emit_copy", "verify"], list(out)[:8]))

    @property
    def launches(self):
        return int(load_library().smash_ctx_launch_count(self.h))

    @property
    def index_bytes(self):
        return int(load_library().smash_ctx_index_bytes(self.h))

    @property
    def stream(self):
        return load_library().smash_ctx_stream(self.h)


class GcNorm:
    """GC normalisation of bin counts on the GPU: cbs.r:18-25 with lowess.gc (cbs.r:3-7), `smash_gcnorm_*`.

    gc_content: gc.txt's gc.content column; chrom_names: its bin.chrom column (chr1..chr22 are the autosomes whose
    mean scales the ratio, cbs.r:13-16, 21).  run(counts) -> (ratio, lowratio); counts may be a host array or a device
    pointer (e.g. the one Context.tail_finish / bins_finish filled)."""

    def __init__(self, gc_content, chrom_names, f=0.05, iter=3, device=0):
        gc = np.ascontiguousarray(gc_content, dtype=np.float64)
        num = [23 if c == "chrX" else 24 if c == "chrY" else int(c[3:]) if c[3:].isdigit() else 99 for c in chrom_names]   # as.numeric -> NA: not an autosome
        auto = np.ascontiguousarray(np.array(num) < 23, dtype=np.uint8)
        self.n = len(gc)
        self.h = C.c_void_p()
        _check(load_library().smash_gcnorm_create(C.c_int(device), _ptr(gc), _ptr(auto), C.c_uint64(self.n), C.c_double(f), C.c_int(iter),
                                                  C.byref(self.h)))

    def run(self, counts=None, counts_device_ptr=None):
        ratio = np.zeros(self.n, dtype=np.float64)
        low = np.zeros(self.n, dtype=np.float64)
        c = np.ascontiguousarray(counts, dtype=np.int64) if counts is not None else None
        _check(load_library().smash_gcnorm_run(self.h, _ptr(c) if c is not None else None,
                                               C.c_void_p(counts_device_ptr) if counts_device_ptr else None, _ptr(ratio), _ptr(low)))
        return ratio, low

    def close(self):
        if self.h:
            load_library().smash_gcnorm_destroy(self.h)
            self.h = C.c_void_p()
