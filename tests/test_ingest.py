"""Input side: raw SAM text / FASTQ pair text -> packed read batch.

CPU: the oracle restatement (oracle/ingest.py) against the golden outputs of the unmodified reference
(tests/golden/case_ingest, made by make_golden_ingest.py), and the host emulation of the kernels' building
blocks (tests/emul/emul_ingest.cpp over csrc/ingest.cuh) against the oracle, including chunked streaming.
GPU (-m gpu): the device-side parse through the C ABI (smash_text_upload / smash_submit_text) against the
oracle, bit-exact arrays, and its SAM output against the reference's golden records."""
import gzip
import os

import numpy as np
import pytest

from emul import emul as E
from helpers import GOLDEN, golden_lines, load_golden_case
from oracle import ingest as I
from oracle import oracle as O
from smash_paper_b200 import synth

ING = os.path.join(GOLDEN, "case_ingest")
KEYS = ("names", "name_off", "seq", "qual", "seq_off", "opt", "opt_off", "read_flag")


def gz(name):
    return gzip.open(os.path.join(ING, name)).read()


def as_batch(d):
    """dict of packed arrays -> the object oracle.Index.map_batch / api.Context.map_batch take."""
    b = synth.ReadBatch(names=d["names"], name_off=d["name_off"], seq=d["seq"], qual=d["qual"], seq_off=d["seq_off"],
                        flags=np.zeros(len(d["read_flag"]), np.uint16), opt=d["opt"], opt_off=d["opt_off"])
    b.read_flag = d["read_flag"]
    return b


def get(d, k):
    return d[k] if isinstance(d, dict) else getattr(d, k)


def assert_same_batch(a, b):
    for k in KEYS:
        x, y = np.asarray(get(a, k)), np.asarray(get(b, k))
        assert x.shape == y.shape and np.array_equal(x, y), k


def concat(parts):
    out = {k: [] for k in KEYS}
    base = dict(name_off=0, seq_off=0, opt_off=0)
    for p in parts:
        for k in ("names", "seq", "qual", "opt", "read_flag"):
            out[k].append(np.asarray(get(p, k)))
        for k in base:
            o = np.asarray(get(p, k))
            out[k].append(o[:-1] + base[k])
            base[k] += int(o[-1])
    res = {k: np.concatenate(v) if v else np.zeros(0) for k, v in out.items()}
    for k in base:
        res[k] = np.concatenate([res[k], [base[k]]]).astype(np.int64)
    return res


def stream(parse, kind, texts, sizes, replace_n=False):
    """Feed the text(s) in chunks the way a streaming host does: unconsumed bytes are passed again."""
    pend = [b"", b""]
    pos = [0, 0]
    parts, n_calls, phase = [], 0, False
    while True:
        for f in range(len(texts)):
            take = sizes[n_calls % len(sizes)]
            pend[f] += texts[f][pos[f]:pos[f] + take]
            pos[f] += take
        final = all(pos[f] >= len(texts[f]) for f in range(len(texts)))
        r = parse(kind, pend[0], pend[1], final, replace_n, phase)
        phase = get(r, "mate2_first_next")
        n_calls += 1
        n = get(r, "n") if isinstance(r, dict) else r.n
        if not final:
            assert n % 2 == 0                       # pairing by arrival parity survives the chunk edge
        parts.append(r)
        cons = get(r, "consumed")
        for f in range(len(texts)):
            pend[f] = pend[f][cons[f]:]
        if final:
            break
        assert n_calls < 100000
    return concat(parts)


def fuzz_sam(rng, n_lines):
    ws = [b"\t", b" ", b"  ", b"\t\t", b" \t ", b"\v", b"\f"]
    out = []
    for i in range(n_lines):
        if rng.random() < 0.08:
            out.append(b"\n")
            continue
        L = int(rng.integers(1, 40))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgtZ", np.uint8), size=L))
        qual = bytes(rng.integers(33, 127, size=L).astype(np.uint8))
        name = b"q%d" % i + [b"", b":0", b":1", b":2", b":"][int(rng.integers(0, 5))]
        flag = [b"0", b"77", b"141", b"64", b"128", b"192", b"+77", b"-1", b"0141", b"4"][int(rng.integers(0, 10))]
        mid = [b"*", b"0", b"0", b"*", b"*", b"0", b"0"]
        if rng.random() < 0.1:
            flag += b"x"; mid = mid[1:]
        opts = [b"X%d:Z:%d" % (j, int(rng.integers(0, 1000))) for j in range(int(rng.integers(0, 4)))]
        sep = lambda: ws[int(rng.integers(0, len(ws)))] if rng.random() < 0.3 else b"\t"   # noqa: E731
        fields = [name, flag] + mid + [seq, qual] + opts
        line = (sep() if rng.random() < 0.1 else b"") + b"".join(f + sep() for f in fields[:-1]) + fields[-1]
        line += [b"", b" ", b"\r", b"\t \r"][int(rng.integers(0, 4))] + b"\n"
        out.append(line)
    text = b"".join(out)
    return text if text.endswith(b"\n") else text + b"\n"


def fuzz_fastq(rng, n_rec, mate):
    out = []
    for i in range(n_rec):
        L = int(rng.integers(0 if rng.random() < 0.05 else 1, 60))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTNn", np.uint8), size=L))
        qual = bytes(rng.integers(33, 127, size=L).astype(np.uint8))
        eol = b"\r\n" if rng.random() < 0.1 and L else b"\n"
        pre = [b"", b"\n", b" \n\n", b"  "][int(rng.integers(0, 4))] if rng.random() < 0.2 else b""
        hdr = b"@r%d" % i + ([b"", b" %d:N:0" % (mate + 1), b"\tsecond third"][int(rng.integers(0, 3))])
        if rng.random() < 0.07:
            out.append(pre + b">" + hdr[1:] + eol + seq + eol)
        else:
            gap = b"\n\n" if rng.random() < 0.1 else b""
            out.append(pre + hdr + eol + seq + eol + gap + b"+" + (b"r%d" % i if rng.random() < 0.3 else b"") + eol + qual + eol)
    return b"".join(out)


# ------------------------------------------------------------------------------------------- CPU

def test_oracle_fastqs_to_sam_matches_reference_binary_output():
    fq1, fq2 = gz("r1.fq.gz"), gz("r2.fq.gz")
    assert I.fastqs_to_sam(fq1, fq2, False) == gz("fastqs_to_sam_0.sam.gz")
    assert I.fastqs_to_sam(fq1, fq2, True) == gz("fastqs_to_sam_1.sam.gz")


@pytest.mark.parametrize("which", ["quirks", "fastq"])
def test_oracle_reader_matches_reference_records(which):
    """The reader restatement feeds the (already pinned) record oracle: the records must be the ones the
    unmodified mummer printed for the same text."""
    case = load_golden_case("case_basic")
    if which == "quirks":
        batch = I.parse_sam_text(gz("quirks.sam.gz"))
    else:
        batch = I.parse_fastq_pair(gz("r1.fq.gz"), gz("r2.fq.gz"), True)
    hdr, lines = golden_lines(os.path.join(ING, f"mapout_{which}.sam.gz"))
    sam = case["oix"].map_batch(as_batch(batch), mode=O.MAM, min_len=20)
    assert sorted(sam.splitlines(keepends=True)) == lines


def test_emulated_kernels_match_oracle_on_golden_inputs():
    q = gz("quirks.sam.gz")
    assert_same_batch(E.ingest(0, q), I.parse_sam_text(q))
    fq1, fq2 = gz("r1.fq.gz"), gz("r2.fq.gz")
    for rep in (False, True):
        assert_same_batch(E.ingest(1, fq1, fq2, True, rep), I.parse_fastq_pair(fq1, fq2, rep))
    sam = gz("fastqs_to_sam_1.sam.gz")
    assert_same_batch(E.ingest(0, sam), I.parse_sam_text(sam))


@pytest.mark.parametrize("seed", range(6))
def test_emulated_kernels_match_oracle_fuzz(seed):
    rng = np.random.default_rng(100 + seed)
    text = fuzz_sam(rng, 300)
    want = I.parse_sam_text(text)
    assert_same_batch(E.ingest(0, text), want)
    for sizes in ([97], [1, 2000, 13], [5000], [len(text) + 5]):
        assert_same_batch(stream(E.ingest, 0, [text], sizes), want)
    n1 = int(rng.integers(50, 120))
    fq1, fq2 = fuzz_fastq(rng, n1, 0), fuzz_fastq(rng, n1 + int(rng.integers(-3, 4)), 1)
    for rep in (False, True):
        want = I.parse_fastq_pair(fq1, fq2, rep)
        assert_same_batch(E.ingest(1, fq1, fq2, True, rep), want)
        for sizes in ([211], [3, 900, 57], [100000]):
            assert_same_batch(stream(E.ingest, 1, [fq1, fq2], sizes, rep), want)


def test_swar_primitives():
    assert E.ingest_lib().emul_ingest_swar_check() == 0


def test_empty_and_ragged_inputs_emulated():
    for text in (b"", b"\n", b"\n\n\n"):
        r = E.ingest(0, text)
        assert r["n"] == 0 and r["consumed"][0] == len(text)
    r = E.ingest(1, b"", b"")
    assert r["n"] == 0
    r = E.ingest(1, b"@a\nAC\n+\nII\n", b"")                     # mate 2 empty: the loop still prints mate 1's record
    assert_same_batch(r, I.parse_fastq_pair(b"@a\nAC\n+\nII\n", b"", False))
    assert r["n"] == 1
    r = E.ingest(0, b"a 0 * 0 0 * * 0 0 ACGT IIII", b"", False)  # non-final, no newline yet: nothing taken
    assert r["n"] == 0 and r["consumed"] == (0, 0)
    one = b"a 77 * 0 0 * * 0 0 ACGT IIII\n"
    r = E.ingest(0, one * 3, b"", False)                          # odd count in a non-final chunk: the last read waits
    assert r["n"] == 2 and r["consumed"][0] == 2 * len(one)


# ---- the kernels of ingest.cu themselves, executed on host threads (tests/emul/cuda_shim) ----------------------

def test_real_kernels_on_host_golden_inputs():
    q = gz("quirks.sam.gz")
    assert_same_batch(E.ingest_kernels(0, q), I.parse_sam_text(q))
    fq1, fq2 = gz("r1.fq.gz"), gz("r2.fq.gz")
    for rep in (False, True):
        assert_same_batch(E.ingest_kernels(1, fq1, fq2, True, rep), I.parse_fastq_pair(fq1, fq2, rep))


def test_real_kernels_on_host_fuzz_and_streaming():
    rng = np.random.default_rng(301)
    text = fuzz_sam(rng, 1500)                                   # 2 tiles of lines, ~7 tiles of chunks
    want = I.parse_sam_text(text)
    assert_same_batch(E.ingest_kernels(0, text), want)
    assert_same_batch(stream(E.ingest_kernels, 0, [text], [40000, 7, 3000]), want)
    fq1, fq2 = fuzz_fastq(rng, 700, 0), fuzz_fastq(rng, 698, 1)
    want = I.parse_fastq_pair(fq1, fq2, True)
    assert_same_batch(E.ingest_kernels(1, fq1, fq2, True, True), want)
    assert_same_batch(stream(E.ingest_kernels, 1, [fq1, fq2], [30000, 11, 9000], True), want)
    for text, code in BAD_SAM[:3]:
        with pytest.raises(E.IngestError) as ei:
            E.ingest_kernels(0, b"ok 0 * 0 0 * * 0 0 AC II\n" * 1100 + text)      # the error sits in the second tile
        assert ei.value.code == code and ei.value.index == 1100


def test_real_kernels_on_host_many_tiles():
    """The same kernels built with scan tiles of 64 x 2 items: 1200 lines / 90 kB of text are 10 tiles of lines and
    ~45 tiles of chunks, and 340 kB in long lines are ~165 tiles, so k_ing_scan_top runs a second round with a carry
    (in the product geometry that takes 16 MB of text)."""
    rng = np.random.default_rng(9)
    text = fuzz_sam(rng, 1200)
    assert_same_batch(E.ingest_kernels_small(0, text), I.parse_sam_text(text))
    fq1, fq2 = fuzz_fastq(rng, 400, 0), fuzz_fastq(rng, 403, 1)
    assert_same_batch(E.ingest_kernels_small(1, fq1, fq2, True, True), I.parse_fastq_pair(fq1, fq2, True))
    lines = []
    for i in range(8):
        L = 21000 + int(rng.integers(0, 999))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=L))
        qual = bytes(rng.integers(33, 127, size=L).astype(np.uint8))
        lines.append(b"long%d\t%d\t*\t0\t0\t*\t*\t0\t0\t" % (i, 77 if i % 2 == 0 else 141) + seq + b"\t" + qual
                     + (b"\tXX:Z:tail  \t YY:i:1" if i % 3 == 0 else b"") + b"\n")
    text = b"".join(lines)
    assert len(text) // 16 // 128 > 128
    assert_same_batch(E.ingest_kernels_small(0, text), I.parse_sam_text(text))


BAD_SAM = [(b"a 77 * 0 0 * * 0 0 ACGT\n", 1), (b"a x77 * 0 0 * * 0 0 ACGT IIII\n", 2), (b"a 99999999999 * 0 0 * * 0 0 ACGT IIII\n", 2),
           (b"a 77 * 0 0 * * 0 0 ACGT III\n", 3), (b"\r\n", 1), (b"ok 0 * 0 0 * * 0 0 AC II\nbad\n", 1)]
BAD_FASTQ = [(b"Xa\nAC\n+\nII\n", 4), (b"@a\nAC\n-\nII\n", 5), (b"@\nAC\n+\nII\n", 6), (b"@a\nAC\n+\n", 7), (b"@a\nAC\n", 7), (b"@a\n", 7),
             (b"@a\nAC GT\n+\nII II\n", 8), (b"@a\nAC\n+\nI\n", 3), (b"@a\n \n+\nI\n", 8)]


def test_rejected_inputs_emulated():
    for text, code in BAD_SAM:
        with pytest.raises(E.IngestError) as ei:
            E.ingest(0, text)
        assert ei.value.code == code, text
        with pytest.raises(I.Undefined):
            I.parse_sam_text(text)
    good = b"@b\nAC\n+\nII\n"
    for text, code in BAD_FASTQ:
        with pytest.raises(E.IngestError) as ei:
            E.ingest(1, text, good)
        assert ei.value.code == code, text
        if code == 8 and b"AC GT" in text:
            continue          # blanks inside the bases line: the reference prints column-shifted SAM (SEQ "AC", QUAL "GT"); rejected by design
        with pytest.raises((I.Undefined, I.RefError)):
            I.parse_fastq_pair(text, good, False)


# ---- the C++ driver's streaming loops, over a fake library (tests/emul/fake_smash.cpp) -------------------------

def _fake_smash():
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
    so = os.path.join(here, "libfake_smash.so")
    srcs = [os.path.join(here, "fake_smash.cpp"), os.path.join(here, "emul_ingest.cpp")]
    deps = srcs + [os.path.join(here, "../../smash_paper_b200/csrc/ingest.cuh"), os.path.join(here, "../../include/smash_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so] + srcs)
    return so


def _dump(batch):
    out = []
    for i in range(batch["n"]):
        no, so, oo = batch["name_off"], batch["seq_off"], batch["opt_off"]
        out.append(bytes(batch["names"][no[i]:no[i + 1]]) + b"\t%d\t" % batch["read_flag"][i] + bytes(batch["seq"][so[i]:so[i + 1]]) + b"\t"
                   + bytes(batch["qual"][so[i]:so[i + 1]]) + bytes(batch["opt"][oo[i]:oo[i + 1]]) + b"\n")
    return b"".join(out)


def _run_driver(workdir, tag, args, files, chunk):
    import glob, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "smash_paper_b200", "bin", "mummer")
    if not os.path.exists(exe):
        pytest.skip("smash_paper_b200/bin/mummer not built")
    d = os.path.join(workdir, "fake_" + tag)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(os.path.join(d, "ref.fa.bin"))
    open(os.path.join(d, "ref.fa"), "w").write(">chr1\nACGT\n")
    open(os.path.join(d, "ref.fa.bin", "rc1.i4.index.bin"), "w").write("present")
    for name, data in files.items():
        open(os.path.join(d, name), "wb").write(data)
    env = dict(os.environ, LD_PRELOAD=_fake_smash(), SMASH_TEXT_CHUNK=str(chunk))
    r = subprocess.run([exe, "-rcref", "-nomap", "-samout"] + args, cwd=d, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = b""
    for f in sorted(glob.glob(os.path.join(d, "mapout", "*.txt")), key=lambda f: int(f.split(".")[-2])):
        data = open(f, "rb").read()
        assert data.startswith(b"@HD\tfake\n")
        got += data[len(b"@HD\tfake\n"):]
    return got


@pytest.mark.parametrize("chunk", [4096, 20000, 1 << 24])
def test_driver_streaming_loops_over_fake_library(workdir, chunk):
    """bin/mummer -samin and -fastqpair with text buffers smaller than the files (and smaller than some lines:
    the buffer has to grow): what reaches the library, chunk after chunk, is the oracle's parse of the whole input."""
    fq1, fq2 = gz("r1.fq.gz"), gz("r2.fq.gz")
    got = _run_driver(workdir, "fq%d" % chunk, ["-fastqpair", "-replaceN", "ref.fa", "r1.fq", "r2.fq"], {"r1.fq": fq1, "r2.fq": fq2}, chunk)
    assert got == _dump(I.parse_fastq_pair(fq1, fq2, True))
    rng = np.random.default_rng(77)
    a, b = fuzz_fastq(rng, 300, 0), fuzz_fastq(rng, 120, 1)          # mate 2 ends early: the rest of mate 1 is never printed
    got = _run_driver(workdir, "fqn%d" % chunk, ["-fastqpair", "ref.fa", "r1.fq", "r2.fq"], {"r1.fq": a, "r2.fq": b}, chunk)
    assert got == _dump(I.parse_fastq_pair(a, b, False))
    got = _run_driver(workdir, "fqm%d" % chunk, ["-fastqpair", "ref.fa", "r1.fq", "r2.fq"], {"r1.fq": b, "r2.fq": a}, chunk)
    assert got == _dump(I.parse_fastq_pair(b, a, False))
    # gzip streams straight in (smash_mapping.sh:19 runs two zcat processes in front of fastqs_to_sam): one gzip member,
    # and a multi-member stream as bgzip / Illumina's bcl2fastq write them; mate 2 stays plain text
    import gzip
    cut = len(a) // 3
    members = gzip.compress(a[:cut]) + gzip.compress(a[cut:])
    for tag, za in (("z1", gzip.compress(a)), ("zm", members)):
        got = _run_driver(workdir, "fq%s%d" % (tag, chunk), ["-fastqpair", "ref.fa", "r1.fq.gz", "r2.fq"], {"r1.fq.gz": za, "r2.fq": b}, chunk)
        assert got == _dump(I.parse_fastq_pair(a, b, False))
    q = gz("quirks.sam.gz")
    long_line = b"big\t77\t*\t0\t0\t*\t*\t0\t0\t" + b"A" * 9000 + b"\t" + b"I" * 9000 + b"\n"
    text = q + long_line + q
    got = _run_driver(workdir, "sam%d" % chunk, ["-samin", "ref.fa", "reads.sam"], {"reads.sam": text}, chunk)
    assert got == _dump(I.parse_sam_text(text))


# ------------------------------------------------------------------------------------------- GPU

@pytest.fixture(scope="module")
def gctx():
    from smash_paper_b200 import api
    case = load_golden_case("case_basic")
    oix = case["oix"]
    ix = api.Index.from_arrays(oix.text, oix.sa, oix.isa, oix.lcp_vec, E.EmulIndex(oix)._raw.reshape(-1), oix.startpos, oix.sizes, oix.descr)
    ctx = api.Context(ix, device=0, min_len=20, nomap=True)
    yield api, ctx, case
    ctx.close(); ix.close()


def gpu_parse(ctx):
    def parse(kind, t0, t1=b"", final=True, replace_n=False, mate2_first=False):
        n, cons = ctx.text_upload(kind, t0, t1, final, replace_n, mate2_first=mate2_first)
        b = ctx.fetch_batch(0)
        assert b.n == n
        b.consumed = cons
        b.mate2_first_next = ctx.mate2_first_next
        return b
    return parse


@pytest.mark.gpu
def test_gpu_parse_matches_oracle_on_golden_inputs(gctx):
    api, ctx, case = gctx
    parse = gpu_parse(ctx)
    q = gz("quirks.sam.gz")
    assert_same_batch(parse(0, q), I.parse_sam_text(q))
    fq1, fq2 = gz("r1.fq.gz"), gz("r2.fq.gz")
    for rep in (False, True):
        assert_same_batch(parse(1, fq1, fq2, True, rep), I.parse_fastq_pair(fq1, fq2, rep))


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["quirks", "fastq"])
def test_gpu_text_to_sam_matches_reference_records(gctx, which):
    """Raw text in, SAM out, all on the device: the records the unmodified fastqs_to_sam | mummer printed."""
    api, ctx, case = gctx
    hdr, lines = golden_lines(os.path.join(ING, f"mapout_{which}.sam.gz"))
    if which == "quirks":
        text = gz("quirks.sam.gz")
        n, _ = ctx.submit_text(0, api.TEXT_SAM, text)
        want = I.parse_sam_text(text)
    else:
        fq1, fq2 = gz("r1.fq.gz"), gz("r2.fq.gz")
        n, _ = ctx.submit_text(0, api.TEXT_FASTQ_PAIR, fq1, fq2, replace_n=True)
        want = I.parse_fastq_pair(fq1, fq2, True)
    res = ctx.wait(0)
    assert n == want["n"] and res.n_reads == n
    assert sorted(res.sam.splitlines(keepends=True)) == lines
    assert res.sam == case["oix"].map_batch(as_batch(want), mode=O.MAM, min_len=20)      # and in input order
    assert res.sam == ctx.map_batch(as_batch(want)).sam                                  # same as the host-parsed path


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_parse_fuzz_and_streaming(gctx, seed):
    api, ctx, case = gctx
    parse = gpu_parse(ctx)
    rng = np.random.default_rng(200 + seed)
    text = fuzz_sam(rng, 3000)                                   # > 1 scan tile of lines, many tiles of chunks
    want = I.parse_sam_text(text)
    assert_same_batch(parse(0, text), want)
    for sizes in ([40000, 7, 3000], [len(text) + 5]):
        assert_same_batch(stream(parse, 0, [text], sizes), want)
    fq1, fq2 = fuzz_fastq(rng, 1500, 0), fuzz_fastq(rng, 1498 + seed, 1)
    for rep in (False, True):
        want = I.parse_fastq_pair(fq1, fq2, rep)
        assert_same_batch(parse(1, fq1, fq2, True, rep), want)
    assert_same_batch(stream(parse, 1, [fq1, fq2], [30000, 11, 9000], True), I.parse_fastq_pair(fq1, fq2, True))


@pytest.mark.gpu
def test_gpu_parse_large_sam_equals_generated_batch(gctx, workdir):
    """200 k reads (60 MB of SAM text): every scan runs over thousands of tiles; the parsed batch must be the
    generator's arrays, and the SAM through the device parse the SAM through the host-parsed batch."""
    api, ctx, case = gctx
    ref = synth.Reference(case["names"], [np.asarray(s, dtype=np.uint8) for s in case["seqs"]])
    reads = synth.make_reads_fast(ref.concat(), 100_000, read_len=150, seed=5)
    path = os.path.join(workdir, "big.sam")
    synth.write_sam(reads, path)
    text = open(path, "rb").read()
    n, cons = ctx.text_upload(api.TEXT_SAM, text)
    assert n == reads.n and cons[0] == len(text)
    got = ctx.fetch_batch(0)
    for k in ("names", "name_off", "seq", "qual", "seq_off"):
        assert np.array_equal(getattr(got, k), getattr(reads, k)), k
    assert np.array_equal(got.read_flag, api.read_flags_from_sam_flags(reads.flags))
    r = ctx.map_resident(api.WANT_SAM)
    sam_dev = ctx.fetch_sam()
    assert int(r.n_reads) == reads.n
    assert sam_dev == ctx.map_batch(reads).sam


@pytest.mark.gpu
def test_gpu_empty_and_rejected_inputs(gctx):
    api, ctx, case = gctx
    for text in (b"", b"\n\n"):
        n, cons = ctx.text_upload(api.TEXT_SAM, text)
        assert n == 0 and cons[0] == len(text)
    n, _ = ctx.submit_text(0, api.TEXT_SAM, b"")
    res = ctx.wait(0)
    assert n == 0 and res.n_reads == 0 and res.sam_bytes == 0
    n, cons = ctx.text_upload(api.TEXT_FASTQ_PAIR, b"@a\nAC\n+\nII\n", b"")
    assert n == 1
    one = b"a 77 * 0 0 * * 0 0 ACGT IIII\n"
    n, cons = ctx.text_upload(api.TEXT_SAM, one * 3, final=False)
    assert n == 2 and cons[0] == 2 * len(one)
    for text, code in BAD_SAM:
        with pytest.raises(api.SmashError, match="error -6"):
            ctx.text_upload(api.TEXT_SAM, text)
    for text, code in BAD_FASTQ:
        with pytest.raises(api.SmashError, match="error -6"):
            ctx.text_upload(api.TEXT_FASTQ_PAIR, text, b"@b\nAC\n+\nII\n")
    q = gz("quirks.sam.gz")                                        # the context still works after rejected inputs
    assert_same_batch(gpu_parse(ctx)(0, q), I.parse_sam_text(q))
