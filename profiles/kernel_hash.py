"""sha256 over the source text of the per-batch kernels an ncu capture describes (+ the device headers they are built
from).  summarize.py stores it next to the DRAM traffic; bench.py recomputes it and drops `roofline.traffic` when the
kernels have changed since the capture.  Adding an unrelated kernel to kernels.cu does not change the hash."""
import hashlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ("k_mam_seed", "k_mam_search", "k_mam_verify", "k_rec_build", "k_rec_xe", "k_sizes", "k_emit_text", "k_emit_copy")


def kernel_source_hash(root=ROOT):
    csrc = os.path.join(root, "smash_paper_b200", "csrc")
    src = open(os.path.join(csrc, "kernels.cu")).read()
    h = hashlib.sha256()
    for k in KERNELS:
        m = re.search(r"^%s\(" % re.escape(k), src, re.M)
        if not m:
            h.update(("missing " + k).encode())
            continue
        end = src.index("\n}\n", m.start())
        h.update(src[m.start():end].encode())
    for fn in ("core.cuh", "records.cuh"):
        h.update(open(os.path.join(csrc, fn), "rb").read())
    return h.hexdigest()


if __name__ == "__main__":
    print(kernel_source_hash())
