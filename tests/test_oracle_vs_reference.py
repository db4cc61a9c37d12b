"""CPU: oracle restatement vs LIVE runs of the unmodified reference binaries (oracle/_ref), fresh
seeds each mode.  Skipped where oracle/_ref is absent."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import tail as T
from smash_paper_b200 import synth

pytestmark = pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def live(workdir):
    d = os.path.join(workdir, "live")
    ref, reads, fa = synth.small_case(d, n_pairs=700, seed=41, n_highcopy=1, highcopy_copies=120)
    O.ref_build_index(fa)
    return dict(dir=d, ref=ref, reads=reads, fa=fa, oix=O.Index.load(fa))


@pytest.mark.parametrize("flags,mode,ml", [([], O.MAM, 20), (["-l", "13"], O.MAM, 13), (["-maxmatch"], O.MEM, 20),
                                           (["-maxmatch", "-l", "15"], O.MEM, 15), (["-mum"], O.MUM, 20), (["-n", "-l", "18"], O.MAM, 18)])
def test_records(live, flags, mode, ml):
    hdr, lines = O.ref_map(live["fa"], os.path.join(live["dir"], "reads.sam"), live["dir"], extra=flags)
    sam = live["oix"].map_batch(live["reads"], mode=mode, min_len=ml, nucleotides_only="-n" in flags, n_threads=4)
    assert hdr == live["oix"].sam_header().encode()
    assert sorted(sam.splitlines(keepends=True)) == lines


def test_index_and_mappability(live):
    o2 = O.Index.build(live["ref"].names, live["ref"].seqs)
    ref = live["oix"]
    assert np.array_equal(o2.text, ref.text) and np.array_equal(o2.sa, ref.sa) and np.array_equal(o2.isa, ref.isa)
    assert np.array_equal(o2.lcp_vec, ref.lcp_vec) and np.array_equal(o2.lcp_m, ref.lcp_m)
    body = np.fromfile(live["fa"] + ".bin/map.bin", dtype=np.uint8)[2:]
    assert np.array_equal(ref.mappability(), body)


def test_tagger(live):
    hdr, lines = O.ref_map(live["fa"], os.path.join(live["dir"], "reads.sam"), live["dir"])
    p = os.path.join(live["dir"], "all.sam")
    open(p, "wb").write(hdr + b"".join(lines))
    body = np.fromfile(live["fa"] + ".bin/map.bin", dtype=np.uint8)[2:]
    mine = T.tag_lines(hdr.splitlines(keepends=True) + lines, live["oix"].descr[::2], live["oix"].sizes[::2], body)
    assert mine == O.ref_mappability_tag(live["fa"], p).splitlines(keepends=True)


def test_chunk_files_are_sorted_by_the_restated_memsam_order(tmp_path):
    """OutputSorter::flush (query.cpp:448-468) sorts a chunk with MemSam::operator< before writing it: every chunk file
    of the unmodified binary must already be in the order of oracle.memsam_sort_key (which the GPU's K5 record_sort and
    its tests use), ties included."""
    from smash_paper_b200 import synth
    d = str(tmp_path)
    ref, reads, fa = synth.small_case(d, n_pairs=600, seed=19)
    O.ref_build_index(fa, mappability=False)
    chunks = O.ref_map_chunks(fa, os.path.join(d, "reads.sam"), d, threads=3)
    key = O.memsam_sort_key(ref.names, ref.sizes)
    assert sum(len(c) for c in chunks) > 1500
    for lines in chunks:
        assert lines == sorted(lines, key=key)
        assert len(set(key(ln) for ln in lines)) == len(lines)          # a strict order: no two lines share a key
