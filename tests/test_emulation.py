"""CPU: the kernels' building blocks (smash_paper_b200/csrc/core.cuh, records.cuh), run lane by lane
on the host by tests/emul, against the golden vectors and the oracle.  No GPU involved; this does
not replace the -m gpu parity tests, it lets kernel logic be debugged in the dev container."""
import gzip
import os

import numpy as np
import pytest

from emul import emul as E
from helpers import golden_lines, golden_variants, load_golden_case, make_case, oracle_tail
from oracle import oracle as O

MAM = [(c, v) for c in ["case_basic", "case_adversarial"] for v in golden_variants(c) if v["mode"] == "mam"]


@pytest.mark.parametrize("case,variant", MAM, ids=[f"{c}-{v['name']}" for c, v in MAM])
@pytest.mark.parametrize("force_exact", [False, True])
def test_anchor_search_and_records_equal_reference(case, variant, force_exact):
    g = load_golden_case(case)
    hdr, lines = golden_lines(variant["path"])
    ex = E.EmulIndex(g["oix"])
    sam, moff, mt, err = ex.map_batch(g["reads"], O.read_flags(g["reads"]), min_len=variant["min_len"],
                                      nuc_only=variant["nuc"], force_exact=force_exact)
    assert sorted(sam.splitlines(keepends=True)) == lines
    # and the same ORDER as the oracle (input order, HI order)
    assert sam == g["oix"].map_batch(g["reads"], min_len=variant["min_len"], nucleotides_only=variant["nuc"])


@pytest.mark.parametrize("seed_k", [6, 9, 12])
def test_seed_length_does_not_change_results(workdir, seed_k):
    ref, reads, fa, oix, body = make_case(os.path.join(workdir, "emul_k"), n_pairs=400, seed=31)
    ex = E.EmulIndex(oix, seed_k=seed_k)
    sam, moff, mt, err = ex.map_batch(reads, O.read_flags(reads), min_len=20)
    osam, ooff, om = oix.map_batch(reads, min_len=20, want_matches=True)
    assert np.array_equal(moff, ooff)
    assert np.array_equal(mt, np.stack([om["ref"], om["query"], om["len"]], 1))
    assert sam == osam


def test_tagged_output(workdir):
    ref, reads, fa, oix, body = make_case(os.path.join(workdir, "emul_tag"), n_pairs=300, seed=33)
    ex = E.EmulIndex(oix, mapbody=body)
    sam, _, _, err = ex.map_batch(reads, O.read_flags(reads), min_len=20, tag_map=True)
    exp = oracle_tail(oix, body, oix.map_batch(reads, min_len=20), os.path.join(workdir, "emul_tag"), fa)
    assert err == 0
    assert sam == b"".join(exp["tagged"])


def test_wordsink_matches_byte_sink():
    """records.cuh WordSink (register-gathered aligned 8-byte stores, byte stores at the region edges)
    writes exactly the bytes the plain sink writes, at every alignment and length."""
    assert E.lib().emul_wordsink_selftest() == 0
