#!/usr/bin/env python3
"""oracle/make_counters.py -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (never linked into the product).

Builds oracle/_ref/mummer-counters and oracle/_ref/mummer-long-counters: the reference's `mummer` with the event
counters of SURVEY.md Appendix D compiled in.  bench.py's CPU leg runs the counting binary on a small sample of the
benchmarked workload to MEASURE the element-granular bytes the reference algorithm touches per read (SURVEY §8d:
B_alg = (E+S)(w+1) + K + Lk*4w + q(1+rec) + 2q) instead of quoting the survey's estimate.

The reference sources are never copied into this repo: they are copied to a scratch directory under /tmp, patched
there by the line-anchored edits below (every edit asserts the text it expects, so a different reference revision
fails loudly), compiled, and only the binaries land in oracle/_ref/ (git-ignored; they travel to the GPU box).
Counters (all per process, relaxed atomics, printed to stderr by a static destructor as one line
`# smash_counters reads=.. calls=.. edge=.. steps=.. lcp=.. linkfail=.. traverse=.. emit=.. links=..`):
  calls/edge  longSA.cpp:327  one top_down_faster call = two edge probes (ref[SA[start]+i], ref[SA[end]+i])
  steps       longSA.cpp:346, :369  one binary-search step = one SA entry + one text byte
  lcp         longSA.h:163, :167  one LCP read of expand_link;  linkfail = expansions that hit the threshold
  traverse    longSA.cpp:405/415/509   emit  longSA.cpp:520 (MAM) / process_match in collectMEMs is not counted
  links       longSA.cpp:524 (MAM) / :384 (suffixlink, MEM)   reads  longSA.cpp:403 (findMEM) / :507 (MAM)
The mapping output of the patched binary is byte-identical to the unpatched one (tests/test_counters.py).
"""
import os
import shutil
import subprocess
import sys
import tempfile

REF = os.environ.get("SMASH_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "mummer.cpp fasta.cpp locked.cpp longSA.cpp memsam.cpp qsufsort.cpp query.cpp util.cpp".split()
FLAGS = "-std=c++11 -Ofast -march=x86-64-v3 -m64 -pthread -w".split()

HEADER = r'''
#ifndef SMASH_COUNTERS_H_
#define SMASH_COUNTERS_H_
#include <cstdio>
struct SmashCounters {
  unsigned long long reads, calls, edge, steps, lcp, linkfail, traverse, emit, links;
  ~SmashCounters() {
    std::fprintf(stderr, "# smash_counters reads=%llu calls=%llu edge=%llu steps=%llu lcp=%llu linkfail=%llu "
                 "traverse=%llu emit=%llu links=%llu\n", reads, calls, edge, steps, lcp, linkfail, traverse, emit, links);
  }
};
extern SmashCounters smash_cnt;
#define SMASH_CNT(f, n) __atomic_fetch_add(&smash_cnt.f, (unsigned long long)(n), __ATOMIC_RELAXED)
#endif
'''

# (file, 1-based line, text that must be on that line, replacement for that text)
EDITS = [
    ("longSA.h", 12, '#include "./size.h"', '#include "./smash_counters.h"\n#include "./size.h"'),
    ("longSA.h", 163, "while (LCP[start] >= link->depth) {",
     "while ((SMASH_CNT(lcp, 1), LCP[start]) >= link->depth) {"),
    ("longSA.h", 164, "if (++exp >= thresh) return false;",
     "if (++exp >= thresh) { SMASH_CNT(linkfail, 1); return false; }"),
    ("longSA.h", 167, "while (end < Nm1 && LCP[end+1] >= link->depth) {",
     "while (end < Nm1 && (SMASH_CNT(lcp, 1), LCP[end+1]) >= link->depth) {"),
    ("longSA.h", 168, "if (++exp >= thresh) return false;",
     "if (++exp >= thresh) { SMASH_CNT(linkfail, 1); return false; }"),
    ("longSA.cpp", 327, "const int64_t cmp_with_first =",
     "SMASH_CNT(calls, 1); SMASH_CNT(edge, 2); const int64_t cmp_with_first ="),
    ("longSA.cpp", 346, "vgl = (int64_t)c", "SMASH_CNT(steps, 1); vgl = (int64_t)c"),
    ("longSA.cpp", 369, "vgl = (int64_t)c", "SMASH_CNT(steps, 1); vgl = (int64_t)c"),
    ("longSA.cpp", 384, "if (m->depth <= 1) {", "SMASH_CNT(links, 1); if (m->depth <= 1) {"),
    ("longSA.cpp", 403, "while (prefix <= P.length()) {", "SMASH_CNT(reads, 1); while (prefix <= P.length()) {"),
    ("longSA.cpp", 405, "traverse(P, prefix, mli, query.min_len);",
     "SMASH_CNT(traverse, 1); traverse(P, prefix, mli, query.min_len);"),
    ("longSA.cpp", 415, "traverse(P, prefix, xmi, P.length());",
     "SMASH_CNT(traverse, 1); traverse(P, prefix, xmi, P.length());"),
    ("longSA.cpp", 507, "while (prefix < P.length()) {", "SMASH_CNT(reads, 1); while (prefix < P.length()) {"),
    ("longSA.cpp", 509, "traverse(P, prefix, cur, P.length());",
     "SMASH_CNT(traverse, 1); traverse(P, prefix, cur, P.length());"),
    ("longSA.cpp", 520, "query.process_match(", "SMASH_CNT(emit, 1); query.process_match("),
    ("longSA.cpp", 524, "cur.depth = cur.depth-1;", "SMASH_CNT(links, 1); cur.depth = cur.depth-1;"),
]


def patch(tmp):
    files = {}
    for name, line, expect, repl in EDITS:
        if name not in files:
            with open(os.path.join(tmp, name)) as f:
                files[name] = f.read().split("\n")
        text = files[name][line - 1]
        if expect not in text:
            raise SystemExit("make_counters: %s:%d does not hold %r (got %r): different reference revision"
                             % (name, line, expect, text))
        files[name][line - 1] = text.replace(expect, repl, 1)
    for name, lines in files.items():
        with open(os.path.join(tmp, name), "w") as f:
            f.write("\n".join(lines))
    with open(os.path.join(tmp, "smash_counters.h"), "w") as f:
        f.write(HEADER)
    with open(os.path.join(tmp, "longSA.cpp"), "a") as f:
        f.write("\nSmashCounters smash_cnt;\n")


def main():
    if not os.path.isdir(REF):
        print("make_counters: %s absent, keeping prebuilt oracle/_ref" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="smash_counters_")
    try:
        for fn in os.listdir(REF):
            if fn.endswith((".cpp", ".h")):
                shutil.copy(os.path.join(REF, fn), tmp)
        patch(tmp)
        cxx = os.environ.get("CXX", "g++")
        for out, defs in (("mummer-counters", []), ("mummer-long-counters", ["-DSINTS", "-DUINTS"])):
            subprocess.check_call([cxx] + FLAGS + defs + ["-o", os.path.join(OUT, out)] + SRC, cwd=tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
