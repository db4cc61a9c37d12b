"""tests/emul/shimlib.py -- TEST INFRASTRUCTURE ONLY.

Builds tests/emul/libsmash_b200_shim.so: the product's own sources (smash_paper_b200/csrc/{api,kernels,tail,mem_search,
ingest,gcnorm}.cu) compiled with g++ against tests/emul/cuda_shim/cuda_runtime.h, so that the CPU test-suite can EXECUTE the
kernels (blocks of OS threads, real barriers, rendezvous-based warp intrinsics) behind the same C ABI and find logic
bugs without a GPU.  The sources are not modified: two purely syntactic rewrites are applied to temporary copies,

    kernel<<<grid, block, smem, stream>>>(args)      ->  shim_bind(kernel, grid, block, smem, stream)(args)
    extern __shared__ <type> name[];                 ->  <type> *name = (<type> *)shim_dynamic_smem();

and the index builder (sabuild.cu, CUB) is replaced by a stub that reports "not available".  The product never loads
this library (smash_paper_b200/api.py loads smash_paper_b200/libsmash_b200.so and fails without a GPU); tests opt in
explicitly through `load()`.
"""
import ctypes as C
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "smash_paper_b200", "csrc")
GEN = os.path.join(HERE, "_shim_src")
SO = os.path.join(HERE, "libsmash_b200_shim.so")
SOURCES = ["api.cu", "kernels.cu", "tail.cu", "mem_search.cu", "ingest.cu", "gcnorm.cu"]

_LAUNCH = re.compile(r"(\b[A-Za-z_]\w*(?:<[^<>;(){}]*>)?)<<<")
_EXTERN_SHARED = re.compile(r"extern\s+__shared__\s+((?:__align__\(\d+\)\s+)?)([A-Za-z_][\w ]*?)\s+(\w+)\[\];")

STUB = r'''
// generated: stands in for sabuild.cu (CUB radix sorts) in the host emulation.  SA / ISA / LCP are canonical functions
// of the text, so a plain comparison sort gives the arrays the kernels downstream expect; tests of the index BUILDER
// itself mean nothing under the emulation and are not run there.
#include "sabuild.cuh"
#include "smash_b200.h"
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <vector>
extern "C" int smash_comm_destroy(smash_ctx *) { return 0; }      // comm.cu (NCCL) is not part of the emulation
namespace smash {
int build_index_device(const uint8_t *T, uint64_t N, int w, void *d_sa, void *d_isa, uint8_t *d_lcp, LcpItem **d_lcpm, uint64_t *n_m,
                       uint64_t, cudaStream_t, char *err, uint64_t *) {
  if (N > (64u << 20)) { if (err) snprintf(err, 200, "host emulation: text too long for the comparison sort"); return -3; }
  std::vector<uint64_t> sa(N), isa(N), lcp(N);
  for (uint64_t i = 0; i < N; ++i) sa[i] = i;
  std::sort(sa.begin(), sa.end(), [&](uint64_t a, uint64_t b) {
    const uint64_t la = N - a, lb = N - b, l = la < lb ? la : lb;
    const int r = memcmp(T + a, T + b, l);
    return r ? r < 0 : la < lb;
  });
  for (uint64_t i = 0; i < N; ++i) isa[sa[i]] = i;
  uint64_t h = 0;
  for (uint64_t i = 0; i < N; ++i) {
    const uint64_t m = isa[i];
    if (m == 0) lcp[m] = 0;
    else { const uint64_t j = sa[m - 1]; while (i + h < N && j + h < N && T[i + h] == T[j + h]) ++h; lcp[m] = h; }
    if (h) --h;
  }
  std::vector<LcpItem> big;
  for (uint64_t i = 0; i < N; ++i) {
    if (w == 4) { ((uint32_t *)d_sa)[i] = (uint32_t)sa[i]; if (d_isa) ((uint32_t *)d_isa)[i] = (uint32_t)isa[i]; }
    else { ((uint64_t *)d_sa)[i] = sa[i]; if (d_isa) ((uint64_t *)d_isa)[i] = isa[i]; }
    d_lcp[i] = (uint8_t)(lcp[i] < 255 ? lcp[i] : 255);
    if (lcp[i] >= 255) big.push_back(LcpItem{i, lcp[i]});
  }
  *n_m = big.size();
  if (cudaMalloc((void **)d_lcpm, sizeof(LcpItem) * (big.size() + 1)) != cudaSuccess) return -4;
  if (!big.empty()) memcpy(*d_lcpm, big.data(), sizeof(LcpItem) * big.size());
  return 0;
}
}
'''


def rewrite(text):
    text = _LAUNCH.sub(lambda m: f"shim_bind({m.group(1)}, ", text)
    text = text.replace(">>>(", ")(")
    text = _EXTERN_SHARED.sub(lambda m: f"{m.group(2)} *{m.group(3)} = ({m.group(2)} *)shim_dynamic_smem();", text)
    text = text.replace('#include "../../include/smash_b200.h"', '#include "smash_b200.h"')
    return text


def build(force=False, sanitize=False):
    """sanitize=True: a second library built with -fsanitize=alignment,bounds (aborts on a misaligned vector load or an
    out-of-range index of a fixed-size array -- both fault or corrupt on the GPU, both are silent on a plain CPU build).
    sanitize="thread": built with ThreadSanitizer (run under LD_PRELOAD=libtsan.so): lanes that exchange data through
    shared or global memory without the __syncwarp / __syncthreads / atomic the GPU needs show up as data races."""
    tag = "" if not sanitize else ("_tsan" if sanitize == "thread" else "_san")
    so = SO.replace(".so", tag + ".so")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps += [os.path.join(HERE, "cuda_shim", "cuda_runtime.h"), os.path.abspath(__file__), os.path.join(ROOT, "include", "smash_b200.h")]
    if not force and os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    os.makedirs(GEN, exist_ok=True)
    objs = []
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            open(os.path.join(GEN, f), "w").write(rewrite(open(os.path.join(CSRC, f)).read()))
    srcs = []
    for f in SOURCES:
        out = os.path.join(GEN, f.replace(".cu", "_shim.cpp"))
        open(out, "w").write(rewrite(open(os.path.join(CSRC, f)).read()))
        srcs.append(out)
    for f in ("compact.h", "expand.h", "expand.cpp", "ctx_internal.h"):     # plain host C++: compiled as it is
        open(os.path.join(GEN, f), "w").write(open(os.path.join(CSRC, f)).read().replace('#include "../../include/smash_b200.h"', '#include "smash_b200.h"'))
    srcs.append(os.path.join(GEN, "expand.cpp"))
    stub = os.path.join(GEN, "sabuild_stub.cpp")
    open(stub, "w").write(STUB)
    srcs.append(stub)
    flags = ["-O1", "-std=c++20", "-fPIC", "-pthread", "-w", "-I", os.path.join(HERE, "cuda_shim"), "-I", GEN, "-I", os.path.join(ROOT, "include")]
    san = [] if not sanitize else (["-fsanitize=thread"] if sanitize == "thread" else ["-fsanitize=alignment,bounds"])
    if sanitize:
        flags += ["-g"] + san + ([] if sanitize == "thread" else ["-fno-sanitize-recover=alignment,bounds"])
    procs = []
    for s in srcs:
        o = s.replace(".cpp", tag + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen(["g++"] + flags + ["-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            raise RuntimeError(f"shim build of {os.path.basename(s)} failed:\n{out[-6000:]}")
    subprocess.check_call(["g++", "-shared", "-pthread"] + san + ["-o", so] + objs)
    return so


def load(sanitize=False):
    return C.CDLL(build(sanitize=sanitize))


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, sanitize="thread" if "--tsan" in sys.argv else "--sanitize" in sys.argv))
