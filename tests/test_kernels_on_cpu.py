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
          "or test_driver_fastq_pair_matches_reference_pipeline or test_saturated_repeat_family_reports_each_match_once")


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


CONTROL = r"""
#include "cuda_runtime.h"
// deliberate bug when sync == 0: lane 0 writes shared memory, the other lanes read it without __syncwarp
__global__ void k_racy(int *out, int sync) {
  __shared__ int box[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) box[warp] = 41 + warp;
  if (sync) __syncwarp();
  out[threadIdx.x] = box[warp];
}
extern "C" int run(int sync) {
  int out[64];
  shim_bind(k_racy, 1, 64, 0, nullptr)(out, sync);
  return out[5];
}
"""


def _tsan():
    try:
        p = subprocess.check_output(["gcc", "-print-file-name=libtsan.so"], text=True).strip()
    except (OSError, subprocess.CalledProcessError):
        return None
    return p if os.path.isabs(p) and os.path.exists(p) else None


def _control(tmp_path):
    shim = os.path.join(ROOT, "tests", "emul", "cuda_shim")
    src = tmp_path / "ctl.cpp"
    src.write_text(CONTROL)
    so = str(tmp_path / "libctl.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++20", "-fPIC", "-shared", "-pthread", "-w", "-fsanitize=thread", "-I", shim, "-o", so, str(src)])
    env = dict(os.environ, LD_PRELOAD=_tsan(), TSAN_OPTIONS="exitcode=0")
    for sync, racy in ((0, True), (1, False)):
        r = subprocess.run([sys.executable, "-c", f"import ctypes; print('ret', ctypes.CDLL({so!r}).run({sync}))"],
                           env=env, capture_output=True, text=True, timeout=300)
        assert "ret 41" in r.stdout, r.stdout + r.stderr
        assert ("ThreadSanitizer: data race" in r.stderr) == racy, r.stderr[-2000:]


@pytest.mark.skipif(_tsan() is None, reason="libtsan not available")
def test_thread_sanitizer_control(tmp_path):
    """Under the emulation ThreadSanitizer reports a deliberately missing __syncwarp and is silent once it is there:
    `run_on_cpu.py --tsan` is a meaningful check of the kernels' synchronisation."""
    _control(tmp_path)


@pytest.mark.skipif(_tsan() is None or os.environ.get("SMASH_TEST_TSAN") != "1",
                    reason="opt-in (SMASH_TEST_TSAN=1): several minutes")
def test_thread_sanitizer_finds_no_race_in_the_kernels():
    """The product's kernels (search, records, SAM emit, tail, MEM mode, MUM mode, input stage) under ThreadSanitizer."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "run_on_cpu.py"), "--tsan", "-x", "-k",
                        "test_mam_matches_and_sam or test_tagged_sam_and_tail or test_gpu_parse_matches_oracle_on_golden_inputs "
                        "or (test_golden_mem_records and basic-mem_l20) or (test_mum_mode and 20)"],
                       capture_output=True, text=True, timeout=7200)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "ThreadSanitizer: data race" not in out, out[-3000:]
