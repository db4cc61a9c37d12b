"""CPU: a fast subset of the `-m gpu` parity tests, run HERE against the product's own kernels executed by the host
emulation (tests/emul/run_on_cpu.py: smash_paper_b200/csrc/*.cu compiled with g++ over tests/emul/cuda_shim).  It
guards kernel logic between GPU runs; it is not a CPU path of the product and not evidence of GPU parity.
The full suite on the CPU: `python tests/emul/run_on_cpu.py` (~15 min)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SUBSET = ("test_header or test_mam_matches_and_sam or test_tagged_sam_and_tail or test_double_buffered_submit "
          "or (test_golden_mam_records and case_basic-mam_l20) or (test_golden_mem_records and case_basic-mem_l20) "
          "or (test_mum_mode and 20) or test_gpu_text_to_sam_matches_reference_records or test_gpu_empty_and_rejected_inputs "
          "or test_driver_fastq_pair_matches_reference_pipeline")


@pytest.mark.skipif(os.environ.get("SMASH_CUDA_SHIM") == "1", reason="already inside the emulation")
def test_gpu_parity_subset_on_emulated_kernels():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "run_on_cpu.py"), "-x", "-k", SUBSET],
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail


@pytest.mark.skipif(os.environ.get("SMASH_CUDA_SHIM") == "1", reason="already inside the emulation")
def test_smoke_on_sanitized_emulated_kernels():
    """The same kernels built with -fsanitize=alignment,bounds: a misaligned vector access faults on the GPU and is
    silent on a plain CPU build; here it aborts the run."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "run_on_cpu.py"), "--sanitize", "-x", "-k",
                        "test_mam_matches_and_sam or test_tagged_sam_and_tail or test_gpu_parse_matches_oracle_on_golden_inputs"],
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
