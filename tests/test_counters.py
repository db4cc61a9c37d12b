"""CPU: the counting build of the reference (oracle/_ref/mummer-counters, oracle/make_counters.py) maps
exactly what the unmodified binary maps, and its counters are the per-read figures bench.py turns into the reference's
algorithmic bytes per read.  Skipped where oracle/_ref holds no counting build."""
import glob
import os
import shutil
import subprocess

import pytest

from oracle import oracle as O
from smash_paper_b200 import synth

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(O.REF_BIN, "mummer-counters")),
                                reason="oracle/_ref/mummer-counters not built")


def run_counting(fa, sam, d, extra=()):
    shutil.rmtree(os.path.join(d, "mapout"), ignore_errors=True)
    p = subprocess.run([os.path.join(O.REF_BIN, "mummer-counters"), "-rcref", "-qthreads", "2", "-nomap", "-samin", "-samout",
                        *extra, fa, sam], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    lines = []
    for fn in glob.glob(os.path.join(d, "mapout", "*.txt")):
        with open(fn, "rb") as f:
            lines += [ln for ln in f if not ln.startswith(b"@")]
    c = [ln for ln in p.stderr.decode().splitlines() if ln.startswith("# smash_counters")]
    assert len(c) == 1
    return sorted(lines), {k: int(v) for k, v in (kv.split("=") for kv in c[0].split()[2:])}


@pytest.mark.parametrize("extra", [(), ("-maxmatch",)])
def test_counting_build_maps_like_the_unmodified_binary(tmp_path, extra):
    d = str(tmp_path)
    ref, reads, fa = synth.small_case(d, n_pairs=400, seed=7)
    O.ref_build_index(fa, mappability=False)
    sam = os.path.join(d, "reads.sam")
    _, want = O.ref_map(fa, sam, d, extra=extra)
    got, c = run_counting(fa, sam, d, extra)
    assert got == want
    assert c["reads"] == 800
    assert c["edge"] == 2 * c["calls"] and c["steps"] > c["calls"] and c["traverse"] >= c["reads"]
    if not extra:
        # MAM: every suffix link is followed by an LCP read unless the depth fell to 0; matches emitted = match records' items
        assert c["links"] > 0 and c["lcp"] >= c["links"] - c["reads"] * 150 and c["emit"] > 0
