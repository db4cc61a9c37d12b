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
