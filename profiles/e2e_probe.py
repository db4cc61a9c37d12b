#!/usr/bin/env python
"""Where the end-to-end time of one batch goes (measurement tooling, not product code).

    SMASH_DEBUG_TIMING=1 python profiles/e2e_probe.py [--reads 1000000] > gpurun_out/e2e_probe.jsonl

config1-sized reference (the transport does not depend on the index size), 1 M-read batches in pinned memory:
per host_threads setting, one isolated smash_map_batch (the library prints its job timeline on stderr) and a pipelined
smash_submit/smash_wait loop over both slots; plus what plain multi-threaded memcpy reaches on this host."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def memcpy_gbs(threads, mb=256, reps=4):
    src = [np.ones(mb << 20, dtype=np.uint8) for _ in range(threads)]
    dst = [np.zeros(mb << 20, dtype=np.uint8) for _ in range(threads)]

    def work(i):
        for _ in range(reps):
            np.copyto(dst[i], src[i])
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    return threads * reps * (mb << 20) / (time.perf_counter() - t0) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--threads", default="1,2,4,8,16")
    ap.add_argument("--slots", default="2")
    a = ap.parse_args()
    import bench
    from smash_paper_b200 import api, sequence, synth
    for t in (1, 4, 16):
        print(json.dumps({"memcpy_threads": t, "gbs_copied": memcpy_gbs(t)}), flush=True)
    wl = bench.WORKLOADS["config1"]
    ref = synth.make_reference(wl["chroms"], seed=1, n_families=wl["families"])
    text, startpos, sizes, descr = sequence.text_from_chromosomes(ref.names, ref.seqs, rcref=True)
    ctx = api.Context.from_text(text, startpos, sizes, descr, keep_isa=True, min_len=20, nomap=True, tag_mappability=True)
    ctx.build_mappability_device(ref.total)
    ctx.drop_isa()
    starts = bench.make_bins(wl, ref, "/tmp")
    ctx.tail_configure(starts, ref.names, ref.offsets())
    genome = ref.concat()
    batches = [bench.pinned_batch(api, synth.make_reads_fast(genome, a.reads // 2, read_len=150, seed=1000, first_pair=i * (a.reads // 2))) for i in range(4)]
    assert max(int(x) for x in a.slots.split(",")) <= min(api.N_SLOTS, 4)
    want = api.WANT_SAM | api.WANT_TAIL
    ctx.tail_reserve((a.reads // 2) * (a.steps + 6), 8 * a.reads * (a.steps + 6))
    for full, mode in ((True, 1), (False, 2), (False, 0)):
        for th, ns in [(t_, n_) for t_ in ([0] if full else [int(x) for x in a.threads.split(",")]) for n_ in [int(x) for x in a.slots.split(",")]]:
            ctx.set_transport(mode=mode, host_threads=th)
            ctx.tail_reset()
            for i in range(2 * ns):
                ctx.submit(i % ns, batches[i % 4], want=want); ctx.wait(i % ns, copy=False)
            print(f"--- mode={mode} threads={th}: isolated batch", file=sys.stderr, flush=True)
            t0 = time.perf_counter()
            ctx.submit(0, batches[0], want=want); r = ctx.wait(0, copy=False)
            one = time.perf_counter() - t0
            print("--- pipelined", file=sys.stderr, flush=True)
            ctx.io_bytes(reset=True)
            t0 = time.perf_counter()
            for i in range(a.steps):
                if i >= ns:
                    ctx.wait(i % ns, copy=False)
                ctx.submit(i % ns, batches[i % 4], want=want)
            for i in range(max(0, a.steps - ns), a.steps):
                ctx.wait(i % ns, copy=False)
            dt = time.perf_counter() - t0
            h2d, d2h = ctx.io_bytes(reset=True)
            print(json.dumps({"transport": {0: "auto", 1: "full", 2: "compact"}[mode], "host_threads": th, "slots": ns, "isolated_batch_ms": one * 1e3, "pipelined_reads_per_s": a.reads * a.steps / dt,
                              "ms_per_batch": dt / a.steps * 1e3, "d2h_bytes_per_batch": d2h // a.steps, "sam_bytes": int(r.sam_bytes)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
