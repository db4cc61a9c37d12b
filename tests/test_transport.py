"""The compact SAM transport (smash_paper_b200/csrc/compact.h): only the text the GPU computes crosses PCIe and host
threads rebuild the lines from the submitted batch.  CPU: the host half (smash_host_expand) against the oracle's lines.
-m gpu: compact and full transport give the same bytes as the oracle in every output mode."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import load_golden_case, make_case, oracle_tail
from oracle import oracle as O


def _read_fields(reads, i):
    name = bytes(reads.names[reads.name_off[i]:reads.name_off[i + 1]])
    seq = bytes(reads.seq[reads.seq_off[i]:reads.seq_off[i + 1]])
    qual = bytes(reads.qual[reads.seq_off[i]:reads.seq_off[i + 1]])
    opt = bytes(reads.opt[reads.opt_off[i]:reads.opt_off[i + 1]]) if reads.opt.size else b""
    return name, seq, qual, opt


def _compact_of(reads, sam):
    """Take the oracle's lines apart the way the device writes them: per record {sam_off, cmp_off, read, head_len,
    tags_len, lr_len | rc << 31} + the text that is not the caller's own bytes.  Lines are in input order."""
    meta = []
    cmp_parts = []
    sam_off = cmp_off = 0
    read = -1
    for line in sam.splitlines(keepends=True):
        f = line[:-1].split(b"\t")
        flag = int(f[1])
        if (flag & 4) or b"HI:i:0" in f:                        # first record of the next read (every read prints at least one)
            read += 1
        name, seq, qual, opt = _read_fields(reads, read)
        assert f[0] == name
        q = len(seq)
        head = b"\t" + b"\t".join(f[1:9]) + b"\t"
        rest = line[len(name) + len(head) + 2 * q + 1:]        # tags [opt] lr-tags \n
        if opt:
            k = rest.index(opt)
            tags, lr = rest[:k], rest[k + len(opt):]
        else:
            k = rest.find(b"\tL0:i:")
            tags, lr = (rest[:k], rest[k:]) if k >= 0 else (rest[:-1], rest[-1:])
        rc = bool(flag & 16)
        meta.append((sam_off, cmp_off, read, len(head), len(tags), len(lr) | (0x80000000 if rc else 0), 0))
        cmp_parts += [head, tags, lr]
        sam_off += len(line)
        cmp_off += len(head) + len(tags) + len(lr)
    m = np.zeros(len(meta), dtype=[("sam_off", "<u8"), ("cmp_off", "<u4"), ("read", "<u4"), ("head", "<u4"), ("tags", "<u4"),
                                   ("lr", "<u4"), ("pad", "<u4")])
    for i, t in enumerate(meta):
        m[i] = t
    return m, b"".join(cmp_parts)


@pytest.mark.parametrize("case", ["case_basic", "case_adversarial"])
def test_host_expand_rebuilds_the_oracles_lines(case):
    from smash_paper_b200 import api
    g = load_golden_case(case)
    reads = g["reads"]
    sam = g["oix"].map_batch(reads, min_len=20, n_threads=4)
    meta, cmp = _compact_of(reads, sam)
    assert meta.itemsize == 32 and len(cmp) < 0.62 * len(sam)
    rf = api._read_flag(reads)
    b = api.Context._cbatch(reads, rf)
    out = np.zeros(len(sam) + 64, dtype=np.uint8)
    cbuf = np.frombuffer(cmp + b"\0" * 64, dtype=np.uint8)
    rc = api.load_library().smash_host_expand(C.byref(b), C.c_uint64(0), meta.ctypes.data_as(C.c_void_p), C.c_uint64(len(meta)),
                                              cbuf.ctypes.data_as(C.c_char_p), out.ctypes.data_as(C.c_char_p))
    assert rc == 0
    assert out[:len(sam)].tobytes() == sam
    assert not out[len(sam):].any()                              # nothing written past the last line
    # record ranges expanded separately, in any order (what the library's host threads do): runs share cache lines
    out2 = np.zeros(len(sam) + 64, dtype=np.uint8)
    cuts = [0, 1, 7, len(meta) // 3, len(meta) // 3 + 1, 2 * len(meta) // 3, len(meta)]
    for a, z in reversed(list(zip(cuts, cuts[1:]))):
        part = meta[a:z]
        assert api.load_library().smash_host_expand(C.byref(b), C.c_uint64(0), part.ctypes.data_as(C.c_void_p), C.c_uint64(len(part)),
                                                    cbuf.ctypes.data_as(C.c_char_p), out2.ctypes.data_as(C.c_char_p)) == 0
    assert out2[:len(sam)].tobytes() == sam and not out2[len(sam):].any()


def test_host_reverse_complement_table():
    """reverse_complement (fasta.cpp:26-61): every byte value, odd lengths (vector body + scalar tail)."""
    from smash_paper_b200 import api, synth
    comp = {ord(a): ord(b) for a, b in zip("acgtrymkbdhvACGTRYMKBDHV", "tgcayrkmvhdbTGCAYRKMVHDB")}
    rng = np.random.default_rng(5)
    for q in (1, 31, 32, 33, 150, 251):
        seq = rng.integers(1, 256, q, dtype=np.uint8)
        seq[seq == 9] = 65; seq[seq == 10] = 67
        qual = rng.integers(33, 127, q, dtype=np.uint8)
        reads = synth.ReadBatch(names=np.frombuffer(b"r", np.uint8), name_off=np.array([0, 1], np.int64), seq=seq, qual=qual,
                                seq_off=np.array([0, q], np.int64), flags=np.array([16], np.uint16), opt=np.zeros(0, np.uint8),
                                opt_off=np.zeros(2, np.int64))
        meta = np.zeros(1, dtype=[("sam_off", "<u8"), ("cmp_off", "<u4"), ("read", "<u4"), ("head", "<u4"), ("tags", "<u4"), ("lr", "<u4"), ("pad", "<u4")])
        meta[0] = (0, 0, 0, 1, 0, 1 | 0x80000000, 0)
        b = api.Context._cbatch(reads, api._read_flag(reads))
        out = np.zeros(2 * q + 4 + 64, dtype=np.uint8)
        cbuf = np.frombuffer(b"\t\n" + b"\0" * 64, dtype=np.uint8)
        assert api.load_library().smash_host_expand(C.byref(b), C.c_uint64(0), meta.ctypes.data_as(C.c_void_p), C.c_uint64(1),
                                                    cbuf.ctypes.data_as(C.c_char_p), out.ctypes.data_as(C.c_char_p)) == 0
        want = b"r\t" + bytes(comp.get(int(c), int(c)) for c in seq[::-1]) + b"\t" + bytes(qual[::-1]) + b"\n"
        assert out[:len(want)].tobytes() == want


# ------------------------------------------------------------------------------------------------ GPU

@pytest.fixture(scope="module")
def case(workdir):
    d = os.path.join(workdir, "transport")
    ref, reads, fa, oix, body = make_case(d, n_pairs=1200, seed=23)
    return dict(dir=d, ref=ref, reads=reads, fa=fa, oix=oix, body=body)


@pytest.mark.gpu
@pytest.mark.parametrize("tagged", [False, True])
@pytest.mark.parametrize("chunks", [1, 4])
def test_full_and_compact_transport_agree(case, tagged, chunks):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, tag_mappability=tagged)
    try:
        sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
        if tagged:
            ctx.load_mappability_file(case["fa"] + ".bin/map.bin")
            sam = b"".join(oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])["tagged"])
        ctx.set_chunking(chunks, 2)
        got = {}
        for mode in (0, 1, 2, 3):                                  # scheduler's choice, full text, compact, ranges alternating
            ctx.set_transport(mode=mode, host_threads=3)
            ctx.io_bytes(reset=True)
            for rep in range(2):
                res = ctx.map_batch(case["reads"])
                assert res.sam == sam, (mode, rep)
            got[mode] = ctx.io_bytes(reset=True)
        assert got[1][0] == got[2][0] == got[0][0]                 # same upload
        assert got[2][1] < 0.62 * got[1][1]                        # ~40 % of the text + 32 B per record come back
        assert got[2][1] <= got[0][1] <= got[1][1]
        if chunks > 1:
            assert got[2][1] < got[3][1] < got[1][1]
    finally:
        ctx.close(); ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["sorted", "mem", "matches", "golden_opt"])
def test_compact_transport_output_modes(case, variant):
    from smash_paper_b200 import api
    if variant == "golden_opt":                                    # optional fields sit between the tags and the L/R tags
        g = load_golden_case("case_adversarial")
        ix = api.Index.open(g["fa"]) if os.path.exists(g["fa"] + ".bin") else None
        if ix is None:
            g["oix"].save(g["fa"])
            ix = api.Index.open(g["fa"])
        reads, oix = g["reads"], g["oix"]
        assert reads.opt.size
    else:
        ix = api.Index.open(case["fa"])
        reads, oix = case["reads"], case["oix"]
    mode = api.MODE_MEM if variant == "mem" else api.MODE_MAM
    ctx = api.Context(ix, min_len=20, nomap=True, mode=mode)
    try:
        want = api.WANT_SAM | (api.WANT_SORTED if variant == "sorted" else 0) | (api.WANT_MATCHES if variant == "matches" else 0)
        out = []
        for mode in (2, 1, 0):
            ctx.set_transport(mode=mode, host_threads=2)
            out.append(ctx.map_batch(reads, want=want).sam)
        assert out[0] == out[1] == out[2]
        exp = oix.map_batch(reads, min_len=20, n_threads=4, mode=O.MEM if variant == "mem" else O.MAM)
        if variant == "sorted":
            key = O.memsam_sort_key(oix.descr[::2], oix.sizes[::2])
            assert out[0] != exp and out[0] == b"".join(sorted(exp.splitlines(keepends=True), key=key))
        else:
            assert out[0] == exp
    finally:
        ctx.close(); ix.close()


@pytest.mark.gpu
def test_workers_keep_submission_order_in_the_tail(case):
    """Two slots, two worker threads: the tail's first-wins dedupe and the positions order follow the order of
    submission; a slot's error arrives at smash_wait; tail calls refuse to run while a batch is in flight."""
    from smash_paper_b200 import api, samio
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, tag_mappability=True)
    try:
        ctx.load_mappability_file(case["fa"] + ".bin/map.bin")
        reads = case["reads"]
        sam = case["oix"].map_batch(reads, min_len=20, n_threads=4)
        exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
        ci = exp["chrominfo"]
        ctx.tail_configure([int(b[2]) for b in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
        ctx.set_chunking(2, 2)
        ctx.set_transport(mode=3, host_threads=2)                  # every batch: one range as full text, one compact
        cuts = [0, 400, 402, 1000, 1700, 1702, reads.n]            # uneven batches (even boundaries: mates stay together)
        parts = [samio.slice_batch(reads, a, b) for a, b in zip(cuts, cuts[1:])]
        for rep in range(2):
            ctx.tail_reset()
            texts = []
            for i, p in enumerate(parts):
                slot = i % api.N_SLOTS
                if i >= api.N_SLOTS:
                    texts.append(api.Result(ctx.wait(slot, copy=False), api.WANT_SAM).sam)
                ctx.submit(slot, p, want=api.WANT_SAM | api.WANT_TAIL)
            with pytest.raises(api.SmashError):
                ctx.tail_finish()                                  # batches still in flight
            for i in range(len(parts) - api.N_SLOTS, len(parts)):
                texts.append(api.Result(ctx.wait(i % api.N_SLOTS, copy=False), api.WANT_SAM).sam)
            assert b"".join(texts) == b"".join(exp["tagged"])
            counts, st = ctx.tail_finish()
            chrom, pos = ctx.tail_positions()
            names = case["oix"].descr[::2]
            assert [f"{names[c]} {p}" for c, p in zip(chrom, pos)] == exp["positions"]
            assert np.array_equal(counts, exp["counts"])
    finally:
        ctx.close(); ix.close()
