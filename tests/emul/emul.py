"""tests/emul/emul.py -- TEST INFRASTRUCTURE ONLY: ctypes driver of the host-side lane-by-lane
emulation (tests/emul/emul.cpp) of the CUDA kernels.  Used by the CPU test-suite to check the
kernels' logic against the oracle when no GPU is present.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libemul.so")
        src = os.path.join(HERE, "emul.cpp")
        deps = [src] + [os.path.join(HERE, "../../smash_paper_b200/csrc", f) for f in ("core.cuh", "records.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
        _LIB = C.CDLL(so)
        _LIB.emul_index_create.restype = C.c_void_p
        _LIB.emul_map_batch.restype = C.c_uint64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class EmulIndex:
    def __init__(self, oix, seed_k=0, mapbody=None):
        """oix: oracle.oracle.Index (only used as a container of the on-disk arrays)."""
        self.oix = oix
        raw = np.zeros((len(oix.lcp_m), 16), dtype=np.uint8)
        if len(oix.lcp_m):
            raw[:, :8] = oix.lcp_m["idx"].astype("<u8").view(np.uint8).reshape(-1, 8)
            raw[:, 8:] = oix.lcp_m["val"].astype("<u8").view(np.uint8).reshape(-1, 8)
        self._raw = raw
        self._descr = (C.c_char_p * len(oix.descr))(*[d.encode() for d in oix.descr])
        self._map = mapbody
        self.h = C.c_void_p(lib().emul_index_create(
            _p(oix.text), C.c_uint64(oix.N), _p(oix.sa), _p(oix.isa), oix.w, _p(oix.lcp_vec), _p(raw),
            C.c_uint64(len(oix.lcp_m)), C.c_uint64(len(oix.descr)), _p(oix.startpos), _p(oix.sizes),
            self._descr, oix.rcref, seed_k, _p(mapbody) if mapbody is not None else None,
            C.c_uint64(0 if mapbody is None else len(mapbody))))

    def map_batch(self, batch, read_flag, min_len=20, nomap=True, nuc_only=False, tag_map=False,
                  force_exact=False):
        n = batch.n
        cap = n * 3000 + 4096
        out = np.empty(cap, dtype=np.uint8)
        moff = np.zeros(n + 1, dtype=np.int64)
        mcap = n * 64 + 64
        mt = np.zeros(3 * mcap, dtype=np.uint64)
        err = C.c_uint32(0)
        need = lib().emul_map_batch(self.h, C.c_uint64(n), _p(batch.names), _p(batch.name_off), _p(batch.seq),
                                    _p(batch.qual), _p(batch.seq_off), _p(batch.opt) if batch.opt.size else None,
                                    _p(batch.opt_off), _p(read_flag), C.c_uint32(min_len), int(nomap),
                                    int(nuc_only), int(tag_map), int(force_exact), _p(out), C.c_uint64(cap),
                                    _p(moff), _p(mt), C.c_uint64(mcap), C.byref(err))
        assert need <= cap and moff[-1] <= mcap
        return out[:need].tobytes(), moff, mt[:3 * moff[-1]].reshape(-1, 3), err.value


# ---- device-side input stage (smash_paper_b200/csrc/ingest.cu) ----------------------------------------
_ING = None


def ingest_lib():
    global _ING
    if _ING is None:
        so = os.path.join(HERE, "libemul_ingest.so")
        src = os.path.join(HERE, "emul_ingest.cpp")
        deps = [src] + [os.path.join(HERE, "../../smash_paper_b200/csrc", f) for f in ("core.cuh", "ingest.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
        _ING = C.CDLL(so)
    return _ING


class IngestError(Exception):
    def __init__(self, code, index):
        super().__init__(f"ingest error {code} at record {index}")
        self.code, self.index = code, index


_INGK = {}


def ingest_kernels_lib(small_tiles=False):
    """ingest.cu itself, compiled with g++ against tests/emul/cuda_shim: its kernels run on host threads.
    small_tiles: scan tiles of 64 x 2 items instead of 256 x 4, so that the multi-round paths of the scans
    (carry between rounds of k_ing_scan_top, many tiles) are reached with small inputs."""
    if small_tiles not in _INGK:
        so = os.path.join(HERE, "libemul_ingest_kernels_small.so" if small_tiles else "libemul_ingest_kernels.so")
        src = os.path.join(HERE, "emul_ingest_kernels.cpp")
        deps = [src, os.path.join(HERE, "cuda_shim", "cuda_runtime.h")] + [os.path.join(HERE, "../../smash_paper_b200/csrc", f)
                                                                          for f in ("core.cuh", "ingest.cuh", "ingest_launch.cuh", "ingest.cu")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas"]
                                  + (["-DING_IB=64", "-DING_II=2"] if small_tiles else [])
                                  + ["-I", os.path.join(HERE, "cuda_shim"), "-o", so, src])
        _INGK[small_tiles] = C.CDLL(so)
    return _INGK[small_tiles]


def ingest_kernels(kind, text0, text1=b"", final=True, replace_n=False, mate2_first=False):
    return ingest(kind, text0, text1, final, replace_n, mate2_first, kernels=True)


def ingest_kernels_small(kind, text0, text1=b"", final=True, replace_n=False, mate2_first=False):
    return ingest(kind, text0, text1, final, replace_n, mate2_first, kernels="small")


def ingest(kind, text0, text1=b"", final=True, replace_n=False, mate2_first=False, kernels=False):
    """Emulated smash_text_upload: dict of the packed batch arrays + consumed bytes.  kernels=False: the building
    blocks of ingest.cuh in plain loops; kernels=True: the kernels of ingest.cu executed block by block."""
    n0, n1 = len(text0), len(text1)
    a0 = np.frombuffer(text0 + b"\0", dtype=np.uint8)
    a1 = np.frombuffer(text1 + b"\0", dtype=np.uint8)
    cap = n0 + n1 + 64
    nl = text0.count(b"\n") + text1.count(b"\n") + 4
    names = np.zeros(cap, np.uint8); seq = np.zeros(cap, np.uint8); qual = np.zeros(cap, np.uint8); opt = np.zeros(cap, np.uint8)
    name_off = np.zeros(nl, np.int64); seq_off = np.zeros(nl, np.int64); opt_off = np.zeros(nl, np.int64)
    rf = np.zeros(nl, np.uint16)
    info = np.zeros(8, np.uint64)
    err_index = C.c_uint64(0)
    fn = ingest_kernels_lib(kernels == "small").emul_ingest_kernels if kernels else ingest_lib().emul_ingest
    rc = fn(int(kind), int(final), int(replace_n), int(mate2_first), _p(a0), C.c_uint64(n0), _p(a1), C.c_uint64(n1),
                                  _p(names), _p(name_off), _p(seq), _p(qual), _p(seq_off), _p(opt), _p(opt_off), _p(rf),
                                  _p(info), C.byref(err_index))
    if rc:
        raise IngestError(rc, err_index.value)
    n, nb, sb, ob = (int(x) for x in info[:4])
    return dict(n=n, names=names[:nb], name_off=name_off[:n + 1], seq=seq[:sb], qual=qual[:sb], seq_off=seq_off[:n + 1],
                opt=opt[:ob], opt_off=opt_off[:n + 1], read_flag=rf[:n], consumed=(int(info[4]), int(info[5])),
                mate2_first_next=bool(info[6]))
