"""-m gpu: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

from helpers import make_case, oracle_tail
from oracle import oracle as O
from oracle import tail as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case(workdir):
    d = os.path.join(workdir, "gpu_small")
    ref, reads, fa, oix, body = make_case(d, n_pairs=1500, seed=11)
    return dict(dir=d, ref=ref, reads=reads, fa=fa, oix=oix, body=body)


@pytest.fixture(scope="module")
def gpu(case):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True)
    yield api, ix, ctx
    ctx.close()
    ix.close()


def _triples(m):
    return np.stack([m["ref"], m["query"], m["len"]], axis=1).astype(np.uint64)


def test_header(case, gpu):
    api, ix, ctx = gpu
    assert ix.sam_header() == case["oix"].sam_header().encode()


def test_mam_matches_and_sam(case, gpu):
    api, ix, ctx = gpu
    sam, moff, mm = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4, want_matches=True)
    res = ctx.map_batch(case["reads"], want=api.WANT_SAM | api.WANT_MATCHES)
    assert np.array_equal(res.match_off, moff)
    assert np.array_equal(res.matches, _triples(mm))
    assert res.sam == sam                       # byte-exact, same (input) order
    assert ctx.launches > 0


@pytest.mark.parametrize("min_len", [16, 24, 31])
def test_min_len_sweep(case, min_len):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=min_len, nomap=True)
    try:
        sam = case["oix"].map_batch(case["reads"], min_len=min_len, n_threads=4)
        res = ctx.map_batch(case["reads"])
        assert res.sam == sam
    finally:
        ctx.close(); ix.close()


def test_no_nomap_and_n_flag(case):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=False, nucleotides_only=True)
    try:
        sam = case["oix"].map_batch(case["reads"], min_len=20, nomap=False, nucleotides_only=True, n_threads=4)
        assert ctx.map_batch(case["reads"]).sam == sam
    finally:
        ctx.close(); ix.close()


def test_mappability_build(case, gpu):
    api, ix, ctx = gpu
    total = int(case["oix"].sizes[::2].sum())
    body = ctx.build_mappability(total)
    assert np.array_equal(body, case["body"])


def test_tagged_sam_and_tail(case):
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, tag_mappability=True)
    try:
        ctx.load_mappability_file(case["fa"] + ".bin/map.bin")
        sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
        exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
        ci = exp["chrominfo"]
        ctx.tail_configure([int(b[2]) for b in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
        res = ctx.map_batch(case["reads"], want=api.WANT_SAM | api.WANT_TAIL)
        assert res.sam == b"".join(exp["tagged"])
        counts, st = ctx.tail_finish()
        chrom, pos = ctx.tail_positions()
        names = case["oix"].descr[::2]
        got = [f"{names[c]} {p}" for c, p in zip(chrom, pos)]
        assert got == exp["positions"]
        assert np.array_equal(counts, exp["counts"])
        assert (st["total_reads"], st["dups_removed"], st["reads_kept"]) == (exp["total"], exp["dups"], exp["kept"])
        assert (st["n_dupe_pairs"], st["n_non_dupe_pairs"]) == (exp["n_dupe"], exp["n_non"])
    finally:
        ctx.close(); ix.close()


@pytest.mark.parametrize("chunks,min_reads", [(4, 2), (3, 2), (2, 100), (4, 10**9)])
def test_chunked_submit_is_byte_identical(case, chunks, min_reads):
    """smash_ctx_set_chunking: a batch cut into read ranges (upload / kernel / download streams)
    gives the same tagged SAM bytes, the same positions and the same bin counts as the whole batch."""
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, tag_mappability=True)
    try:
        ctx.load_mappability_file(case["fa"] + ".bin/map.bin")
        sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
        exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
        ci = exp["chrominfo"]
        ctx.tail_configure([int(b[2]) for b in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
        ctx.set_chunking(chunks, min_reads)
        for rep in range(2):                                   # second pass: buffers already sized
            ctx.tail_reset()
            ctx.submit(0, case["reads"], want=api.WANT_SAM | api.WANT_TAIL)
            res = ctx.wait(0)
            assert res.sam == b"".join(exp["tagged"])
            counts, st = ctx.tail_finish()
            chrom, pos = ctx.tail_positions()
            names = case["oix"].descr[::2]
            assert [f"{names[c]} {p}" for c, p in zip(chrom, pos)] == exp["positions"]
            assert np.array_equal(counts, exp["counts"])
            assert (st["total_reads"], st["dups_removed"], st["reads_kept"]) == (exp["total"], exp["dups"], exp["kept"])
        with pytest.raises(api.SmashError):
            ctx.set_chunking(9, 2)
    finally:
        ctx.close(); ix.close()


def test_double_buffered_submit(case, gpu):
    api, ix, ctx = gpu
    sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
    ctx.submit(0, case["reads"]); ctx.submit(1, case["reads"])
    a = ctx.wait(0); b = ctx.wait(1)
    assert a.sam == sam and b.sam == sam


def _lcpm_raw(oix):
    raw = np.zeros((len(oix.lcp_m), 16), dtype=np.uint8)
    if len(oix.lcp_m):
        raw[:, :8] = oix.lcp_m["idx"].astype("<u8").view(np.uint8).reshape(-1, 8)
        raw[:, 8:] = oix.lcp_m["val"].astype("<u8").view(np.uint8).reshape(-1, 8)
        if oix.w == 4:
            raw[:, 12:] = 0
    return raw.reshape(-1)


@pytest.mark.parametrize("chunk_cap", [0, 30000])
def test_gpu_index_build_matches_canonical_arrays(case, chunk_cap):
    """SA/ISA/LCP built on the GPU == the arrays of the index files (canonical functions of the text)."""
    from smash_paper_b200 import api
    oix = case["oix"]
    ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=4, chunk_cap=chunk_cap)
    try:
        sa, isa, vec, m = ctx.copy_index(oix.N, 4)
        assert np.array_equal(sa, oix.sa)
        assert np.array_equal(isa, oix.isa)
        assert np.array_equal(vec, oix.lcp_vec)
        assert np.array_equal(m, _lcpm_raw(oix))
        assert ctx.map_batch(case["reads"]).sam == oix.map_batch(case["reads"], min_len=20, n_threads=4)
    finally:
        ctx.close()


def test_gpu_index_files_equal_reference_files(case, workdir):
    """smash_ctx_save_index writes byte-identical <fa>.bin/ files (incl. map.bin body)."""
    import filecmp, shutil
    from smash_paper_b200 import api
    oix = case["oix"]
    d = os.path.join(workdir, "saved")
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "ref.fa")
    shutil.copy(case["fa"], fa)
    ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=4)
    try:
        ctx.build_mappability(int(oix.sizes[::2].sum()))
        ctx.save_index(fa, with_mappability=True)
    finally:
        ctx.close()
    for f in ["rc1.ref.bin", "rc1.ref.seq.bin", "rc1.i4.index.bin", "rc1.i4.index.sa.bin", "rc1.i4.index.isa.bin",
              "rc1.i4.index.lcp.vec.bin", "rc1.i4.index.lcp.m.bin"]:
        assert filecmp.cmp(os.path.join(fa + ".bin", f), os.path.join(case["fa"] + ".bin", f), shallow=False), f
    a = np.fromfile(fa + ".bin/map.bin", dtype=np.uint8)[2:]
    assert np.array_equal(a, case["body"])


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built")
def test_against_unmodified_reference_binary(workdir):
    """End to end against the real `mummer` (oracle/_ref) at 2 x 1 Mb: index files, map.bin, SAM lines."""
    from smash_paper_b200 import api, synth
    d = os.path.join(workdir, "vs_ref")
    os.makedirs(d, exist_ok=True)
    ref = synth.make_reference([("chr1", 1000000), ("chr2", 800000)], seed=5, n_families=30, n_long=3)
    fa = os.path.join(d, "ref.fa")
    synth.write_fasta(ref, fa)
    reads = synth.make_reads(ref, 5000, seed=6)
    synth.write_sam(reads, os.path.join(d, "reads.sam"))
    O.ref_build_index(fa)
    hdr, lines = O.ref_map(fa, os.path.join(d, "reads.sam"), d, threads=4)
    ix = api.Index.open(fa)
    ctx = api.Context(ix, min_len=20, nomap=True)
    try:
        assert ix.sam_header() == hdr
        res = ctx.map_batch(reads)
        assert sorted(res.sam.splitlines(keepends=True)) == lines
        body = ctx.build_mappability(ref.total)
        assert np.array_equal(body, np.fromfile(fa + ".bin/map.bin", dtype=np.uint8)[2:])
    finally:
        ctx.close(); ix.close()
    # and the GPU-built index equals the reference-built files
    oix = O.Index.load(fa)
    ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=4)
    try:
        sa, isa, vec, m = ctx.copy_index(oix.N, 4)
        assert np.array_equal(sa, oix.sa) and np.array_equal(isa, oix.isa) and np.array_equal(vec, oix.lcp_vec)
        assert np.array_equal(m, _lcpm_raw(oix))
    finally:
        ctx.close()


# ---- golden vectors of the unmodified reference (tests/golden) through the CUDA path ----------
from helpers import golden_lines, golden_variants, load_golden_case  # noqa: E402

_GOLD_MAM = [(c, v) for c in ["case_basic", "case_adversarial"] for v in golden_variants(c) if v["mode"] == "mam"]


@pytest.mark.parametrize("case_name,variant", _GOLD_MAM, ids=[f"{c}-{v['name']}" for c, v in _GOLD_MAM])
def test_golden_mam_records(case_name, variant):
    from smash_paper_b200 import api
    g = load_golden_case(case_name)
    hdr, lines = golden_lines(variant["path"])
    oix = g["oix"]
    ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=4, min_len=variant["min_len"],
                                nomap=True, nucleotides_only=variant["nuc"])
    try:
        res = ctx.map_batch(g["reads"])
        assert sorted(res.sam.splitlines(keepends=True)) == lines                 # what the reference printed
        assert res.sam == oix.map_batch(g["reads"], min_len=variant["min_len"], nucleotides_only=variant["nuc"], n_threads=4)
    finally:
        ctx.close()


def test_golden_map_bin():
    import gzip
    from smash_paper_b200 import api
    for case_name in ["case_basic", "case_adversarial"]:
        g = load_golden_case(case_name)
        oix = g["oix"]
        ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=4)
        try:
            body = ctx.build_mappability(int(oix.sizes[::2].sum()))
            ref = np.frombuffer(gzip.open(os.path.join(g["dir"], "map.bin.gz")).read(), dtype=np.uint8)[2:]
            assert np.array_equal(body, ref)
        finally:
            ctx.close()


@pytest.mark.parametrize("driver", ["cxx", "cxx_host_reader", "python"])
def test_mummer_compatible_driver(case, workdir, driver):
    """The drop-in `mummer` (C++ binary over the C ABI, and its Python twin): same flags, same files;
    index built on the GPU when absent."""
    import shutil, subprocess, sys, glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = [os.path.join(root, "smash_paper_b200", "bin", "mummer")] if driver.startswith("cxx") else [sys.executable, "-m", "smash_paper_b200.mummer"]
    d = os.path.join(workdir, "driver_" + driver)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    fa = os.path.join(d, "ref.fa")
    shutil.copy(case["fa"], fa)
    shutil.copy(os.path.join(case["dir"], "reads.sam"), os.path.join(d, "reads.sam"))
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if driver == "cxx_host_reader":
        env["SMASH_HOST_READER"] = "1"                   # -samin parsed by the host QueryParser instead of ingest.cu
    # index build (index_setup.sh:19): exits 1 on 'dummy' by design, leaves <fa>.bin/ behind
    r = subprocess.run(exe + ["-verbose", "-rcref", fa, "dummy"], cwd=d, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 1 and "unable to open dummy" in r.stderr
    import filecmp
    for f in ["rc1.ref.bin", "rc1.ref.seq.bin", "rc1.i4.index.bin", "rc1.i4.index.sa.bin", "rc1.i4.index.isa.bin",
              "rc1.i4.index.lcp.vec.bin", "rc1.i4.index.lcp.m.bin"]:
        assert filecmp.cmp(os.path.join(fa + ".bin", f), os.path.join(case["fa"] + ".bin", f), shallow=False), f
    r = subprocess.run(exe + ["-rcref", "-mappability", fa, fa + ".bin/map.bin"],
                       cwd=d, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(fa + ".bin/map.bin", dtype=np.uint8)[2:], case["body"])
    r = subprocess.run(exe + ["-rcref", "-qthreads", "12", "-nomap", "-samin", "-samout",
                        fa, "reads.sam"], cwd=d, env=dict(env, SMASH_CHUNK_ORDER="input"), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = b"".join(open(f, "rb").read() for f in sorted(glob.glob(os.path.join(d, "mapout", "*.txt"))))
    exp = case["oix"].sam_header().encode() + case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
    assert got == exp
    # default: a chunk file is what a reference worker writes for the same chunk -- header + lines in MemSam order
    shutil.rmtree(os.path.join(d, "mapout"))
    r = subprocess.run(exe + ["-rcref", "-qthreads", "12", "-nomap", "-samin", "-samout", fa, "reads.sam"], cwd=d, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = b"".join(open(f, "rb").read() for f in sorted(glob.glob(os.path.join(d, "mapout", "*.txt"))))
    hdr = case["oix"].sam_header().encode()
    key = O.memsam_sort_key(case["ref"].names, case["ref"].sizes)
    assert got == hdr + b"".join(sorted(exp[len(hdr):].splitlines(keepends=True), key=key))
    r = subprocess.run(exe + ["-rcref", "-nomap", fa, "reads.sam"], cwd=d, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Error\n-nomap can only be used with -sam_out")


def test_driver_streams_sam_text_in_chunks(case, workdir):
    """`mummer -samin` with the reader on the GPU and a text buffer far smaller than the file: chunk edges cut
    lines and pairs, the concatenated chunk files must still be the single-batch SAM."""
    import shutil, subprocess, glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "smash_paper_b200", "bin", "mummer")
    d = os.path.join(workdir, "driver_chunks")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    env = dict(os.environ, SMASH_TEXT_CHUNK="30011", SMASH_CHUNK_ORDER="input")
    r = subprocess.run([exe, "-rcref", "-nomap", "-samin", "-samout", case["fa"], os.path.join(case["dir"], "reads.sam")],
                       cwd=d, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    files = sorted(glob.glob(os.path.join(d, "mapout", "*.txt")), key=lambda f: int(f.split(".")[-2]))
    assert len(files) > 3
    hdr = case["oix"].sam_header().encode()
    got = b""
    for f in files:
        data = open(f, "rb").read()
        assert data.startswith(hdr)
        got += data[len(hdr):]
    assert got == case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)


@pytest.mark.parametrize("gzipped", [False, True])
def test_driver_fastq_pair_matches_reference_pipeline(workdir, gzipped):
    """`mummer -fastqpair -replaceN ref r1.fq r2.fq` (both FASTQ files parsed on the GPU, streamed in small chunks)
    against what the unmodified `fastqs_to_sam r1.fq r2.fq 1 | mummer -samin` printed (tests/golden/case_ingest);
    gzipped: the .fq.gz files go in as they are (smash_mapping.sh:19 puts a zcat in front of each)."""
    import shutil, subprocess, glob, gzip
    from helpers import GOLDEN, golden_lines
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "smash_paper_b200", "bin", "mummer")
    d = os.path.join(workdir, "driver_fastqpair" + ("_gz" if gzipped else ""))
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    r1, r2 = ("r1.fq.gz", "r2.fq.gz") if gzipped else ("r1.fq", "r2.fq")
    for name, src in (("ref.fa", "case_basic/ref.fa.gz"), (r1, "case_ingest/r1.fq.gz"), (r2, "case_ingest/r2.fq.gz")):
        raw = open(os.path.join(GOLDEN, src), "rb").read()
        open(os.path.join(d, name), "wb").write(raw if name.endswith(".gz") else gzip.decompress(raw))
    env = dict(os.environ, SMASH_TEXT_CHUNK="9000")
    r = subprocess.run([exe, "-rcref", "-nomap", "-samout", "-fastqpair", "-replaceN", "ref.fa", r1, r2],
                       cwd=d, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hdr, lines = golden_lines(os.path.join(GOLDEN, "case_ingest", "mapout_fastq.sam.gz"))
    got = []
    files = glob.glob(os.path.join(d, "mapout", "*.txt"))
    assert len(files) > 3
    for f in files:
        data = open(f, "rb").read()
        assert data.startswith(hdr)
        got += data[len(hdr):].splitlines(keepends=True)
    assert sorted(got) == lines


def test_saturated_repeat_family_reports_each_match_once(workdir):
    """Found by fuzzing the emulated kernels: a read that runs through a > 255 bp repeat family (U saturated) makes
    EVERY copy's candidate fall back to exact_start at the same query start, so the one true match was staged several
    times; the rank sort then left slots unwritten and match_cnt counted the copies (bogus zero-length records).
    The match CSR must equal the oracle's per read -- counts included."""
    from smash_paper_b200 import api
    kw = dict(n_chrom=2, chrom_len=4098, n_pairs=51, seed=945792, read_len=400, n_families=4, family_len=229, n_long=2,
              n_highcopy=0, highcopy_copies=42, n_pad=200, sub_rate=0.0, z_rate=0.01)
    ref, reads, fa, oix, body = make_case(os.path.join(workdir, "satrep"), **kw)
    ix = api.Index.open(fa)
    for min_len, nuc in ((9, True), (12, False), (20, False)):
        ctx = api.Context(ix, device=0, min_len=min_len, nomap=True, nucleotides_only=nuc)
        try:
            res = ctx.map_batch(reads, want=api.WANT_SAM | api.WANT_MATCHES)
            for i in range(reads.n):
                q = bytes(reads.seq[reads.seq_off[i]:reads.seq_off[i + 1]]).lower()
                if nuc:
                    q = bytes(c if c in b"acgt" else ord("~") for c in q)
                want = [tuple(int(x) for x in m) for m in oix.mam(q, min_len)]
                got = [tuple(int(x) for x in m) for m in res.matches[res.match_off[i]:res.match_off[i + 1]]]
                assert got == want, (min_len, nuc, i)
            assert res.sam == oix.map_batch(reads, min_len=min_len, nucleotides_only=nuc)
        finally:
            ctx.close()
    ix.close()


def test_two_shards_equal_one_run(case):
    """Read-sharded tail on ONE GPU: two contexts take the two halves of the pairs; with the key and edge
    exchange of multigpu.py (done in-process here) the summed counts equal the single-context run."""
    import torch
    from smash_paper_b200 import api, multigpu, samio
    ix = api.Index.open(case["fa"])
    base = case["reads"]
    cut = (base.n // 4) * 2
    # second shard = second half of the reads + a copy of 200 reads of the first shard (cross-shard dupes)
    extra = samio.slice_batch(base, 100, 300)
    extra.names = extra.names.copy()
    extra.names[extra.name_off[:-1]] = ord("x")          # unique names that still sort after r*: same keys, later pairs
    reads = samio.concat_batches([base, extra])
    exp = None
    ctxs = []
    try:
        for _ in range(3):
            c = api.Context(ix, min_len=20, nomap=True)
            c.load_mappability_file(case["fa"] + ".bin/map.bin")
            sam = case["oix"].map_batch(reads, min_len=20, n_threads=4)
            if exp is None:
                exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
            ci = exp["chrominfo"]
            c.tail_configure([int(b[2]) for b in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
            ctxs.append(c)
        whole, a, b = ctxs
        whole.map_batch(reads, want=api.WANT_TAIL)
        counts, st = whole.tail_finish()
        assert np.array_equal(counts, exp["counts"])
        half = cut
        a.map_batch(samio.slice_batch(reads, 0, half), want=api.WANT_TAIL)
        b.map_batch(samio.slice_batch(reads, half, reads.n), want=api.WANT_TAIL)
        dev = torch.device("cuda", 0)
        ba, bb = multigpu.ContextBackend(a, 0, dev), multigpu.ContextBackend(b, half // 2, dev)
        ka, kb = ba.export_keys(), bb.export_keys()
        ea = ba.phase_a(ka[:0]); eb = bb.phase_a(multigpu.lower_rank_keys([ka, kb], 1))
        ca, sa = ba.phase_b(False, 0)
        hp, pv = multigpu.previous_last_pos([ea, eb], 1)
        cb, sb = bb.phase_b(hp, pv)
        torch.cuda.synchronize()
        assert np.array_equal((ca + cb).cpu().numpy(), exp["counts"])
        for k in multigpu.STAT_KEYS:
            assert sa[k] + sb[k] == st[k], k
        assert st["n_dupe_pairs"] == exp["n_dupe"] and st["n_non_dupe_pairs"] == exp["n_non"]
    finally:
        for c in ctxs:
            c.close()
        ix.close()


@pytest.mark.parametrize("tagged", [False, True])
def test_record_sort_equals_outputsorter_order(case, tagged):
    """SMASH_WANT_SORTED: the batch's lines in the order OutputSorter::flush writes a chunk (MemSam::operator<,
    memsam.h:136-158; the restated key is pinned on the reference's own chunk files in test_oracle_vs_reference.py),
    same bytes as the input-order output, and the tail is unaffected."""
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, tag_mappability=tagged)
    try:
        if tagged:
            ctx.load_mappability_file(case["fa"] + ".bin/map.bin")
        key = O.memsam_sort_key(case["ref"].names, case["ref"].sizes)
        plain = ctx.map_batch(case["reads"], want=api.WANT_SAM).sam
        for rep in range(2):
            got = ctx.map_batch(case["reads"], want=api.WANT_SAM | api.WANT_SORTED).sam
            assert got == b"".join(sorted(plain.splitlines(keepends=True), key=key))
        ctx.submit(0, case["reads"], want=api.WANT_SAM | api.WANT_SORTED); ctx.submit(1, case["reads"], want=api.WANT_SAM)
        assert ctx.wait(0).sam == got and ctx.wait(1).sam == plain
        # two reads with one name and the same mate bits at one position: the reference throws "flags equal"
        from smash_paper_b200 import samio
        twice = samio.concat_batches([samio.slice_batch(case["reads"], 0, 40), samio.slice_batch(case["reads"], 0, 40)])
        with pytest.raises(api.SmashError, match="flags equal"):
            ctx.map_batch(twice, want=api.WANT_SAM | api.WANT_SORTED)
        assert ctx.map_batch(case["reads"], want=api.WANT_SAM | api.WANT_SORTED).sam == got        # the context survives it
    finally:
        ctx.close(); ix.close()


def _perm_pairs(reads, order):
    from smash_paper_b200 import samio
    return samio.concat_batches([samio.slice_batch(reads, 2 * int(p), 2 * int(p) + 2) for p in order])


@pytest.mark.parametrize("case_name", ["case_basic", "case_tail"])
@pytest.mark.parametrize("feed", ["one_batch", "three_batches", "shuffled_input", "name_sorted_input"])
def test_golden_smashmem_stage(case_name, feed):
    """The fused tail against what the UNMODIFIED smashMEM.py + awk/perl + varbin.py printed (tests/golden/
    make_golden_smash.py).  case_tail's read names are NOT in `samtools sort -n` order: the positions list must come out
    in name order whatever order (and however many batches) the reads arrive in."""
    import gzip
    from smash_paper_b200 import api, samio
    g = load_golden_case(case_name)
    d, oix = g["dir"], g["oix"]
    smash = gzip.open(os.path.join(d, "smash.txt.gz")).read().splitlines()
    gold_pos = gzip.open(os.path.join(d, "positions.txt.gz")).read().decode().splitlines()
    gold_counts = np.array([int(r.split("\t")[3]) for r in gzip.open(os.path.join(d, "varbin.txt.gz")).read().decode().splitlines()])
    nd, nn = (int(x.split()[0]) for x in smash[-1].decode().split("\t"))
    bins = T.read_table(os.path.join(d, "bins.txt"))
    ci = T.read_chrominfo(os.path.join(d, "chrom_sizes.txt"))
    reads = g["reads"]
    n_pairs = reads.n // 2
    rng = np.random.default_rng(3)
    if feed == "shuffled_input":
        reads = _perm_pairs(reads, rng.permutation(n_pairs))
    elif feed == "name_sorted_input":
        import functools
        nm = [bytes(reads.names[reads.name_off[2 * p]:reads.name_off[2 * p + 1]]) for p in range(n_pairs)]
        reads = _perm_pairs(reads, sorted(range(n_pairs), key=functools.cmp_to_key(lambda x, y: T.strnum_cmp(nm[x], nm[y]) or x - y)))
    ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=4, min_len=20, nomap=True, tag_mappability=True)
    try:
        ctx.build_mappability_device()
        ctx.tail_configure([int(b[2]) for b in bins], list(ci.keys()), [int(v[2]) for v in ci.values()])
        if feed == "three_batches":
            cuts = [0, 2 * (n_pairs // 3), 2 * (2 * n_pairs // 3) , reads.n]
            sam = b"".join(ctx.map_batch(samio.slice_batch(reads, a, b), want=api.WANT_SAM | api.WANT_TAIL).sam for a, b in zip(cuts, cuts[1:]))
        else:
            sam = ctx.map_batch(reads, want=api.WANT_SAM | api.WANT_TAIL).sam
        tagged = [ln for ln in gzip.open(os.path.join(d, "tagged.sam.gz")).read().splitlines(keepends=True) if not ln.startswith(b"@")]
        assert sorted(sam.splitlines(keepends=True)) == sorted(tagged)            # mappability_tag binary's output
        counts, st = ctx.tail_finish()
        chrom, pos = ctx.tail_positions()
        names = oix.descr[::2]
        assert [f"{names[c]} {p}" for c, p in zip(chrom, pos)] == gold_pos        # smashMEM.py | awk | perl
        assert np.array_equal(counts, gold_counts)                                 # varbin.py
        assert (st["n_dupe_pairs"], st["n_non_dupe_pairs"]) == (nd, nn)           # smashMEM.py trailer
        assert st["n_positions"] == len(gold_pos) and st["reads_kept"] == int(gold_counts.sum())
    finally:
        ctx.close()


_GOLD_MEM = [(c, v) for c in ["case_basic", "case_adversarial"] for v in golden_variants(c) if v["mode"] == "mem"]


@pytest.mark.parametrize("case_name,variant", _GOLD_MEM, ids=[f"{c}-{v['name']}" for c, v in _GOLD_MEM])
def test_golden_mem_records(case_name, variant):
    """-maxmatch (longSA::MEM): faithful control flow incl. the prefix=1 quirk, the expand_link threshold
    and libstdc++ sort ties -- byte-exact vs what the reference printed."""
    from smash_paper_b200 import api
    g = load_golden_case(case_name)
    hdr, lines = golden_lines(variant["path"])
    oix = g["oix"]
    ctx = api.Context.from_text(oix.text, oix.startpos, oix.sizes, oix.descr, w=4, mode=api.MODE_MEM,
                                min_len=variant["min_len"], nomap=True)
    try:
        res = ctx.map_batch(g["reads"], want=api.WANT_SAM | api.WANT_MATCHES)
        assert sorted(res.sam.splitlines(keepends=True)) == lines
        osam, ooff, om = oix.map_batch(g["reads"], mode=O.MEM, min_len=variant["min_len"], n_threads=4, want_matches=True)
        assert np.array_equal(res.match_off, ooff)
        assert np.array_equal(res.matches, _triples(om))                      # emission order too
        assert res.sam == osam
    finally:
        ctx.close()


@pytest.mark.parametrize("seed_k", [5, 7, 9, 11])
@pytest.mark.parametrize("mode", ["mam", "mum"])
def test_seed_length_prefilter_and_split_search(case, seed_k, mode):
    """Short seeds put many chance candidates into every bucket: the 4+4 character pre-filter is built
    (N / 4^k >= 0.25), k_mam_search parks the survivors and k_mam_verify extends them (split search),
    rows overflow into the in-kernel path, big buckets take the exact path -- same bytes as the oracle."""
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True, seed_k=seed_k, mode=api.MODE_MUM if mode == "mum" else api.MODE_MAM)
    try:
        exp = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4, mode=O.MUM if mode == "mum" else O.MAM)
        assert ctx.map_batch(case["reads"]).sam == exp
    finally:
        ctx.close(); ix.close()


@pytest.mark.parametrize("min_len", [20, 14])
def test_mum_mode(case, min_len):
    """-mum (longSA::MUM, longSA.cpp:549-585): MAM + cleanMUMcand sweep, survivors in by_ref order."""
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, mode=api.MODE_MUM, min_len=min_len, nomap=True)
    try:
        osam, ooff, om = case["oix"].map_batch(case["reads"], mode=O.MUM, min_len=min_len, n_threads=4, want_matches=True)
        res = ctx.map_batch(case["reads"], want=api.WANT_SAM | api.WANT_MATCHES)
        assert np.array_equal(res.match_off, ooff) and np.array_equal(res.matches, _triples(om))
        assert res.sam == osam
    finally:
        ctx.close(); ix.close()
    if O.have_reference():
        hdr, lines = O.ref_map(case["fa"], os.path.join(case["dir"], "reads.sam"), case["dir"], extra=["-mum", "-l", str(min_len)])
        assert sorted(osam.splitlines(keepends=True)) == lines


def test_large_sample_parity_and_invariants():
    """10 Mb reference, 120k reads: byte-exact SAM vs the oracle on the GPU-built index, plus the
    size-independent properties used at BASELINE.json's full sizes: batch-split invariance of SAM and
    counts, sum(counts) == reads kept, records == SAM lines, NH/HI consistency."""
    from smash_paper_b200 import api, samio, sequence, synth
    ref = synth.make_reference([(f"chr{i + 1}", 2_500_000) for i in range(4)], seed=17, n_families=60, n_long=4)
    text, startpos, sizes, descr = sequence.text_from_chromosomes(ref.names, ref.seqs)
    reads = synth.make_reads_fast(ref.concat(), 60_000, seed=5)
    ctx = api.Context.from_text(text, startpos, sizes, descr, min_len=20, nomap=True)
    try:
        sa, isa, vec, m = ctx.copy_index(len(text), 4)
        lcpm = np.zeros(len(m) // 16, dtype=O.LCPM_DT)
        if len(m):
            r = m.reshape(-1, 16)
            lcpm["idx"] = r[:, :8].copy().view("<u8").reshape(-1); lcpm["val"] = r[:, 8:12].copy().view("<u4").reshape(-1)
        oix = O.Index(text, sa, isa, vec, lcpm, startpos, sizes, descr, 1, 4)
        exp = oix.map_batch(reads, min_len=20, n_threads=os.cpu_count() or 4)
        ctx.build_mappability(ref.total)
        starts = np.arange(0, ref.total, 50_000, dtype=np.int64)
        ctx.tail_configure(starts, ref.names, ref.offsets())
        res = ctx.map_batch(reads, want=api.WANT_SAM | api.WANT_TAIL)
        assert res.sam == exp
        counts, st = ctx.tail_finish()
        assert counts.sum() == st["reads_kept"] and st["total_reads"] == st["reads_kept"] + st["dups_removed"]
        assert res.sam.count(b"\n") == sum(1 for _ in exp.splitlines())
        # the same reads in 5 uneven batches: identical SAM bytes and identical counts
        ctx.tail_reset()
        cuts = [0, 10_000, 10_002, 55_554, 90_000, reads.n]
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            parts.append(ctx.map_batch(samio.slice_batch(reads, a, b), want=api.WANT_SAM | api.WANT_TAIL, first_pair=a // 2).sam)
        assert b"".join(parts) == exp
        counts2, st2 = ctx.tail_finish()
        assert np.array_equal(counts, counts2) and st == st2
        # short seed: pre-filter + split search at scale (rows of 32 parked candidates overflow here)
        ctx2 = api.Context.from_text(text, startpos, sizes, descr, min_len=20, nomap=True, seed_k=10)
        try:
            assert ctx2.map_batch(reads).sam == exp
        finally:
            ctx2.close()
        # NH == number of records of the read, HI = 0..NH-1 in order
        import re
        by_read = {}
        for ln in exp.splitlines()[:50_000]:
            f = ln.split(b"\t")
            mm = re.search(rb"NH:i:(\d+)(?:\tHI:i:(\d+))?", ln)
            by_read.setdefault((f[0], int(f[1]) & 192), []).append((int(mm.group(1)), int(mm.group(2)) if mm.group(2) else 0))
        for recs in list(by_read.values())[:-1]:
            if recs[0][0] == 0:
                assert len(recs) == 1
            else:
                assert [h for _, h in recs] == list(range(len(recs))) and all(n == len(recs) for n, _ in recs)
    finally:
        ctx.close()


@pytest.mark.parametrize("read_len,min_len", [(100, 16), (250, 24), (151, 20), (36, 20)])
def test_read_length_and_min_len_sweep(workdir, read_len, min_len):
    """BASELINE.json config 5: 100-250 bp reads, min MEM 16-24 (search divergence / anchor stride changes)."""
    from smash_paper_b200 import api
    d = os.path.join(workdir, f"sweep_{read_len}_{min_len}")
    ref, reads, fa, oix, body = make_case(d, n_pairs=500, seed=read_len + min_len, read_len=read_len, frag_min=2, frag_max=6)
    ix = api.Index.open(fa)
    ctx = api.Context(ix, min_len=min_len, nomap=True, tag_mappability=True)
    try:
        ctx.load_mappability(body)
        sam = oix.map_batch(reads, min_len=min_len, n_threads=4)
        exp = oracle_tail(oix, body, sam, d, fa)
        assert ctx.map_batch(reads).sam == b"".join(exp["tagged"])
    finally:
        ctx.close(); ix.close()


@pytest.mark.parametrize("parts", [1, 2, 10])
def test_bin_resolution_sweep(case, workdir, parts):
    """BASELINE.json config 4: finer bin sets (the 100k/500k bins.txt are missing from the reference
    checkout, so they are split from the coarse set); covers both histogram kernels."""
    from smash_paper_b200 import api, synth
    src = os.path.join(case["dir"], "bins.txt")
    dst = os.path.join(workdir, f"bins_x{parts}.txt")
    synth.split_bins(src, dst, parts)
    bins = T.read_table(dst)
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True)
    try:
        ctx.load_mappability(case["body"])
        sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
        exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
        counts_exp, total, dups, kept = T.varbin(exp["positions"], bins, exp["chrominfo"])
        ci = exp["chrominfo"]
        ctx.tail_configure([int(b[2]) for b in bins], list(ci.keys()), [int(v[2]) for v in ci.values()])
        ctx.map_batch(case["reads"], want=api.WANT_TAIL)
        counts, st = ctx.tail_finish()
        assert np.array_equal(counts, np.array(counts_exp))
        assert (st["total_reads"], st["dups_removed"], st["reads_kept"]) == (total, dups, kept)
    finally:
        ctx.close(); ix.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["late_first_bin", "one_bin", "tiny_bins", "unsorted", "dense_at_the_end"])
def test_bin_lookup_equals_bisect_on_odd_bin_tables(case, workdir, kind):
    """varbin.py:89-92 is `bisect.bisect(starts, abspos) - 1` with python's negative index for positions before the first
    start.  The device lookup narrows the search with a granule table (tail.cu bin_of): bin tables that stress it --
    first bin far from 0 (counts[-1] case), one bin, thousands of bins inside one granule, bins piled up at the end, and an
    unsorted table (plain bisect, whatever it answers) -- must count exactly like python's bisect on the same positions."""
    from smash_paper_b200 import api
    ix = api.Index.open(case["fa"])
    ctx = api.Context(ix, min_len=20, nomap=True)
    try:
        ctx.load_mappability(case["body"])
        sam = case["oix"].map_batch(case["reads"], min_len=20, n_threads=4)
        exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
        ci = exp["chrominfo"]
        total = sum(int(v[1]) for v in ci.values())
        rng = np.random.default_rng(11)
        if kind == "late_first_bin":
            starts = list(range(total // 3, total, 997))
        elif kind == "one_bin":
            starts = [total // 2]
        elif kind == "tiny_bins":
            starts = sorted(set([0] + [int(x) for x in rng.integers(total // 4, total // 4 + 3000, 2500)] + [total // 2]))
        elif kind == "dense_at_the_end":
            starts = [0, 10] + list(range(total - 5000, total, 3))
        else:
            starts = [int(x) for x in rng.permutation(np.arange(0, total, 1009))]
        ctx.tail_configure(starts, list(ci.keys()), [int(v[2]) for v in ci.values()])
        ctx.map_batch(case["reads"], want=api.WANT_TAIL)
        counts, st = ctx.tail_finish()
        # the restated varbin.py loop (oracle/tail.py, pinned on the reference script) on the oracle's positions list
        want, n_total, n_dups, n_kept = T.varbin(exp["positions"], [["c", "0", str(x)] for x in starts], ci)
        assert n_total > 500
        assert (st["total_reads"], st["dups_removed"], st["reads_kept"]) == (n_total, n_dups, n_kept)
        assert np.array_equal(counts, np.array(want))
    finally:
        ctx.close(); ix.close()


def test_error_paths(workdir, case):
    """Errors the reference raises as paa::Error come back as SmashError with the same wording."""
    import shutil
    from smash_paper_b200 import api
    d = os.path.join(workdir, "errors")
    shutil.rmtree(d, ignore_errors=True)
    shutil.copytree(case["dir"], d)
    fa = os.path.join(d, "ref.fa")
    with pytest.raises(api.SmashError, match="could not open reference bin file"):
        api.Index.open(os.path.join(d, "reads.sam"))                    # no <file>.bin/ next to it
    with open(fa, "ab") as f:
        f.write(b"\n")                                                   # FASTA size guard (fasta.cpp:115-119)
    with pytest.raises(api.SmashError, match="reference fasta size has changed"):
        api.Index.open(fa)
    ix = api.Index.open(case["fa"])
    try:
        with pytest.raises(api.SmashError, match="mode"):
            api.Context(ix, mode=7)
        ctx = api.Context(ix, min_len=20, nomap=True)
        with pytest.raises(api.SmashError, match="map.bin"):
            ctx.tail_configure([0, 10], ["chr1"], [0])                  # tail without mappability
        ctx.close()
    finally:
        ix.close()


def test_three_shards_with_verdict_equal_one_run(case):
    """The many-rank form of the sharded tail on ONE GPU: three contexts, the global first-wins verdict
    computed from all exported fingerprints (what multigpu.partitioned_min_ordinals produces with its
    all_to_all exchange), smash_tail_phase_a_verdict + edges + phase_b; summed counts == single run."""
    import torch
    from smash_paper_b200 import api, multigpu, samio
    ix = api.Index.open(case["fa"])
    base = case["reads"]
    extra = samio.slice_batch(base, 40, 440)
    extra.names = extra.names.copy()
    extra.names[extra.name_off[:-1]] = ord("x")
    reads = samio.concat_batches([base, extra])
    sam = case["oix"].map_batch(reads, min_len=20, n_threads=4)
    exp = oracle_tail(case["oix"], case["body"], sam, case["dir"], case["fa"])
    ci = exp["chrominfo"]
    cuts = [0, (reads.n // 6) * 2, (reads.n // 3) * 2 + 200, reads.n]
    ctxs, backs = [], []
    try:
        dev = torch.device("cuda", 0)
        for a, b in zip(cuts[:-1], cuts[1:]):
            c = api.Context(ix, min_len=20, nomap=True)
            c.load_mappability(case["body"])
            c.tail_configure([int(x[2]) for x in exp["bins"]], list(ci.keys()), [int(v[2]) for v in ci.values()])
            c.map_batch(samio.slice_batch(reads, a, b), want=api.WANT_TAIL)
            ctxs.append(c); backs.append(multigpu.ContextBackend(c, a // 2, dev))
        keys = [bk.export_keys() for bk in backs]
        allk = torch.cat(keys, 0)
        uniq, inv = torch.unique(allk[:, :2], dim=0, return_inverse=True)
        gmin = torch.full((uniq.shape[0],), 2 ** 62, dtype=torch.int64, device=dev).scatter_reduce(0, inv, allk[:, 2], "amin")
        mins = gmin[inv]
        edges, o = [], 0
        for bk, k in zip(backs, keys):
            edges.append(bk.phase_a_verdict(mins[o:o + k.shape[0]])); o += k.shape[0]
        total = torch.zeros(len(exp["bins"]), dtype=torch.int64, device=dev)
        stats = {k: 0 for k in multigpu.STAT_KEYS}
        for r, bk in enumerate(backs):
            hp, pv = multigpu.previous_last_pos(edges, r)
            cts, st = bk.phase_b(hp, pv)
            total += cts
            for k in multigpu.STAT_KEYS:
                stats[k] += st[k]
        torch.cuda.synchronize()
        assert np.array_equal(total.cpu().numpy(), exp["counts"])
        assert (stats["total_reads"], stats["dups_removed"], stats["reads_kept"]) == (exp["total"], exp["dups"], exp["kept"])
        assert (stats["n_dupe_pairs"], stats["n_non_dupe_pairs"]) == (exp["n_dupe"], exp["n_non"]) and exp["n_dupe"] >= 200
    finally:
        for c in ctxs:
            c.close()
        ix.close()


def test_long_reads(workdir):
    """Reads longer than the shared-memory staging buffer (1024) take the HBM-staged exact search path."""
    from smash_paper_b200 import api
    d = os.path.join(workdir, "long_reads")
    ref, reads, fa, oix, body = make_case(d, n_pairs=60, seed=77, read_len=1500, frag_min=10, frag_max=30)
    ix = api.Index.open(fa)
    ctx = api.Context(ix, min_len=20, nomap=True)
    try:
        sam = oix.map_batch(reads, min_len=20, n_threads=4)
        assert ctx.map_batch(reads).sam == sam
    finally:
        ctx.close(); ix.close()


@pytest.mark.parametrize("mode", ["mam", "mum"])
def test_more_matches_per_read_than_the_stage_holds(workdir, mode):
    """The reference keeps any number of MAMs per read.  4000-base reads made of 120..160 short unique pieces give ~100 MAMs
    each, more than the anchor kernels stage in shared memory (64): the range is redone by the exact per-start kernel
    with CSR slots (k_mam_exact, count + write passes) and the SAM must still equal the oracle's byte for byte; a second
    batch of ordinary reads on the same context goes back to the anchor path."""
    from smash_paper_b200 import api
    d = os.path.join(workdir, "many_matches_" + mode)
    ref, reads, fa, oix, body = make_case(d, n_pairs=12, seed=91, read_len=4000, frag_min=120, frag_max=160)
    ix = api.Index.open(fa)
    ctx = api.Context(ix, min_len=20, nomap=True, mode=api.MODE_MUM if mode == "mum" else api.MODE_MAM)
    try:
        omode = O.MUM if mode == "mum" else O.MAM
        sam = oix.map_batch(reads, min_len=20, n_threads=4, mode=omode)
        res = ctx.map_batch(reads, want=api.WANT_SAM | api.WANT_MATCHES)
        assert res.sam == sam
        per_read = np.diff(res.match_off)
        assert per_read.max() > 64, per_read.max()
        from smash_paper_b200 import synth
        short = synth.make_reads(ref, 50, read_len=150, seed=5)
        assert ctx.map_batch(short).sam == oix.map_batch(short, min_len=20, n_threads=4, mode=omode)
    finally:
        ctx.close(); ix.close()


def _driver_case_tail(workdir, tag):
    """case_tail of tests/golden on disk: ref.fa + index (built by the driver), map.bin, reads.sam, bins, chrom sizes."""
    import gzip, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "smash_paper_b200", "bin", "mummer")
    src = os.path.join(root, "tests", "golden", "case_tail")
    d = os.path.join(workdir, "driver_bins_" + tag)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    fa = os.path.join(d, "ref.fa")
    open(fa, "wb").write(gzip.open(os.path.join(src, "ref.fa.gz")).read())
    open(os.path.join(d, "reads.sam"), "wb").write(gzip.open(os.path.join(src, "reads.sam.gz")).read())
    shutil.copy(os.path.join(src, "bins.txt"), d); shutil.copy(os.path.join(src, "chrom_sizes.txt"), d)
    r = subprocess.run([exe, "-rcref", fa, "dummy"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "unable to open dummy" in r.stderr
    r = subprocess.run([exe, "-rcref", "-mappability", fa, fa + ".bin/map.bin"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(fa + ".bin/map.bin", "rb").read()[2:] == gzip.open(os.path.join(src, "map.bin.gz")).read()[2:]
    return exe, d, fa, src


@pytest.mark.parametrize("order", ["as_given", "name_sorted"])
@pytest.mark.parametrize("gpus", [1, 2])
def test_driver_fused_tail_equals_reference_varbin(workdir, gpus, order):
    """bin/mummer -bins ..: mapout + the stages after mummer (mappability_tag, smashMEM.py, chromosome filter, varbin.py)
    in one run.  The varbin file must be BYTE-identical to what the unmodified pipeline printed for these reads
    (tests/golden/case_tail/varbin.txt.gz, incl. Python's repr of the ratio column); with -gpus 2 the reads are cut into
    two ranges, each GPU maps its own, and the counts meet in smash_bins_finish (one ncclAllReduce).
    The reference name-sorts the records before smashMEM.py (smash_mapping.sh:23), so its counts do not depend on the
    order of the input: "as_given" (97 pairs out of `samtools sort -n` order: on 2 GPUs the shards are gathered on rank 0,
    which sorts by name) and "name_sorted" (the read-sharded exchange proper) must both print the golden file."""
    import glob, gzip, subprocess
    from smash_paper_b200 import api
    if api.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    exe, d, fa, src = _driver_case_tail(workdir, f"{gpus}_{order}")
    if order == "name_sorted":
        import functools
        otail = T
        lines = open(os.path.join(d, "reads.sam"), "rb").read().splitlines(keepends=True)
        pairs = [(lines[i], lines[i + 1]) for i in range(0, len(lines), 2)]
        pairs.sort(key=functools.cmp_to_key(lambda a, b: otail.strnum_cmp(a[0].split(b"\t")[0], b[0].split(b"\t")[0])))
        open(os.path.join(d, "reads.sam"), "wb").write(b"".join(a + b for a, b in pairs))
    cmd = [exe, "-rcref", "-qthreads", "4", "-nomap", "-samin", "-samout", "-bins", "bins.txt", "-chromsizes", "chrom_sizes.txt",
           "-binout", "varbin.txt", "-binstats", "stats.txt"] + (["-gpus", str(gpus)] if gpus > 1 else []) + [fa, "reads.sam"]
    r = subprocess.run(cmd, cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(os.path.join(d, "varbin.txt"), "rb").read() == gzip.open(os.path.join(src, "varbin.txt.gz")).read()
    stats = open(os.path.join(d, "stats.txt")).read().splitlines()
    assert stats[0] == "TotalReads\tDupsRemoved\tReadsKept\tMedianBinCount" and len(stats[1].split("\t")) == 4
    # the mapout of all chunk files = the reference's records for these reads (sorted multiset; chunking differs by design)
    got = sorted(l for f in glob.glob(os.path.join(d, "mapout", "*.txt")) for l in open(f, "rb") if not l.startswith(b"@"))
    g = load_golden_case("case_tail") if os.path.exists(os.path.join(src, "mapout_mam_l20.sam.gz")) else None
    if g is None:
        from smash_paper_b200 import samio, sequence
        names, seqs = sequence.read_fasta(fa)
        oix = O.Index.build(names, seqs)
        exp = sorted(oix.map_batch(samio.read_sam(os.path.join(d, "reads.sam")), min_len=20, n_threads=4).splitlines(keepends=True))
        assert got == exp
    if gpus > 1:
        assert len({os.path.basename(f).split("_")[1] for f in glob.glob(os.path.join(d, "mapout", "*.txt"))}) == gpus


def test_driver_gc_normalisation(workdir):
    """bin/mummer -bins .. -gc gc.txt -gcout lowratio.txt: the counts go on to the head of cbs.r (cbs.r:18-25, lowess.gc)
    on the GPU; ratio and lowratio per bin against the oracle's restatement of R's lowess/approx (tests/test_gcnorm.py pins
    it) on the counts the same run printed."""
    import subprocess
    from oracle import gcnorm as G
    exe, d, fa, src = _driver_case_tail(workdir, "gc")
    bins = [l.split("\t") for l in open(os.path.join(d, "bins.txt")).read().splitlines()]
    rng = np.random.default_rng(3)
    gc = np.round(rng.uniform(0.32, 0.6, len(bins)), 6)
    with open(os.path.join(d, "gc.txt"), "w") as f:
        f.write("bin.chrom\tbin.start\tbin.end\tbin.length\tgene.count\tcgi.count\tdist.telomere\tgc.content\tcviki_count\tnlaiii_count\n")
        for b, g in zip(bins, gc):
            f.write("%s\t%s\t0\t0\t0\t0\t0\t%.6f\t0\t0\n" % (b[0], b[1], g))
    cmd = [exe, "-rcref", "-qthreads", "4", "-nomap", "-samin", "-samout", "-bins", "bins.txt", "-chromsizes", "chrom_sizes.txt",
           "-binout", "varbin.txt", "-gc", "gc.txt", "-gcout", "lowratio.txt", fa, "reads.sam"]
    r = subprocess.run(cmd, cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    counts = np.array([int(l.split("\t")[3]) for l in open(os.path.join(d, "varbin.txt")).read().splitlines()])
    rows = [l.split("\t") for l in open(os.path.join(d, "lowratio.txt")).read().splitlines()]
    assert rows[0] == ["chrom", "chrompos", "abspos", "bincount", "ratio", "gc.content", "lowratio"] and len(rows) == len(bins) + 1
    assert [int(x[3]) for x in rows[1:]] == list(counts)
    oratio, olow = G.gc_normalise(counts, gc, [b[0] for b in bins])
    assert np.allclose([float(x[4]) for x in rows[1:]], oratio, rtol=1e-13, atol=0)
    assert np.allclose([float(x[6]) for x in rows[1:]], olow, rtol=1e-9, atol=0)
    r = subprocess.run(cmd[:-6] + ["-gc", "gc.txt", fa, "reads.sam"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "-gc and -gcout go together" in r.stderr
