#!/usr/bin/env python
"""tests/emul/run_on_cpu.py -- TEST INFRASTRUCTURE ONLY.

Runs the `-m gpu` parity tests HERE, without a GPU, against the product's own kernels executed by the host emulation
(tests/emul/shimlib.py builds smash_paper_b200/csrc/*.cu with g++ over tests/emul/cuda_shim/cuda_runtime.h).

    python tests/emul/run_on_cpu.py                       # 56 of the 62 parity tests (~15 min; see BUILDER_TESTS for the rest)
    python tests/emul/run_on_cpu.py -k "golden and mam"   # any pytest selection
    python tests/emul/run_on_cpu.py --sanitize -k golden  # kernels built with -fsanitize=alignment,bounds
    python tests/emul/run_on_cpu.py --tsan -k test_header # ThreadSanitizer: a missing __syncwarp/__syncthreads is a data race

What it is for: finding logic, indexing, scan, shuffle and synchronisation bugs in a kernel change BEFORE spending
GPU time on it.  What it is not: a CPU path of the product (nothing under smash_paper_b200/ can load it; the library
it builds lives under tests/emul and is git-ignored), a performance tool, or evidence of GPU parity -- the `-m gpu`
run on a B200 remains the parity gate.  Tests of the index builder (sabuild.cu, CUB) are excluded: the emulation
replaces it with a comparison sort.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

# not run under the emulation: the index BUILDER is stubbed (these would only test the stub); two tests hand torch CUDA
# tensors to the library; two are sized for a GPU (10 Mb reference / 200 k reads: > 10 min each on OS threads)
BUILDER_TESTS = [
    "tests/test_gpu_parity.py::test_gpu_index_build_matches_canonical_arrays",
    "tests/test_gpu_parity.py::test_gpu_index_files_equal_reference_files",
    "tests/test_gpu_parity.py::test_two_shards_equal_one_run",
    "tests/test_gpu_parity.py::test_three_shards_with_verdict_equal_one_run",
    "tests/test_gpu_parity.py::test_large_sample_parity_and_invariants",
    "tests/test_ingest.py::test_gpu_parse_large_sam_equals_generated_batch",
    "tests/test_gcnorm.py::test_gpu_gc_normalise_small_tied_and_device_counts",     # hands a torch CUDA tensor to the library
]


def env_for_shim(sanitize=False):
    import shimlib
    so = shimlib.build(sanitize=sanitize)
    env = dict(os.environ)
    env["SMASH_B200_LIB"] = so                                   # smash_paper_b200/api.py (tests opt in explicitly)
    env["LD_PRELOAD"] = so + (":" + env["LD_PRELOAD"] if env.get("LD_PRELOAD") else "")    # smash_paper_b200/bin/mummer
    env["SMASH_CUDA_SHIM"] = "1"
    if sanitize == "thread":
        tsan = subprocess.check_output(["gcc", "-print-file-name=libtsan.so"], text=True).strip()
        env["LD_PRELOAD"] = tsan + ":" + env["LD_PRELOAD"]
        env.setdefault("TSAN_OPTIONS", "halt_on_error=0 report_signal_unsafe=0 exitcode=66")
    return env


def main(argv):
    sanitize = "thread" if "--tsan" in argv else "--sanitize" in argv    # sanitizer builds of the kernels
    argv = [a for a in argv if a not in ("--sanitize", "--tsan")]
    args = [sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-q", "-p", "no:cacheprovider", "--timeout=900"]
    for t in BUILDER_TESTS:
        args += ["--deselect", t]
    return subprocess.call(args + argv, cwd=ROOT, env=env_for_shim(sanitize))


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
