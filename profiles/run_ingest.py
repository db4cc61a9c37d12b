#!/usr/bin/env python
"""The command profiled for the input stage (ingest.cu): 1 M reads of SAM text (330 MB, what fastqs_to_sam pipes into
`mummer -samin`) parsed on the GPU, `--calls` times; optionally the same reads as a FASTQ pair.

    python profiles/run_ingest.py [--pairs 500000] [--calls 3] [--fastq]

Prints one JSON line with CUDA-event device time (end of the H2D copy -> end of k_ing_copy) and wall time per call.
The index is a 200 kb toy (the input stage never touches it)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from smash_paper_b200 import api, sequence, synth  # noqa: E402


def fastq_texts(batch):
    """The batch as two FASTQ texts (4 fixed-width lines per record), built with array operations."""
    n, ln, q = batch.n, int(batch.name_off[1]), int(batch.seq_off[1])
    names = batch.names.reshape(n // 2, 2, ln); seq = batch.seq.reshape(n // 2, 2, q); qual = batch.qual.reshape(n // 2, 2, q)
    out = []
    for m in (0, 1):
        w = 1 + ln + 1 + q + 1 + 2 + q + 1
        row = np.empty((n // 2, w), dtype=np.uint8)
        o = 0
        row[:, o] = ord("@"); o += 1
        row[:, o:o + ln] = names[:, m]; o += ln
        row[:, o] = 10; o += 1
        row[:, o:o + q] = np.where(seq[:, m] == ord("Z"), ord("N"), seq[:, m]); o += q
        row[:, o] = 10; o += 1
        row[:, o] = ord("+"); row[:, o + 1] = 10; o += 2
        row[:, o:o + q] = qual[:, m]; o += q
        row[:, o] = 10
        out.append(row.reshape(-1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=500_000)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--fastq", action="store_true")
    args = ap.parse_args()
    ref = synth.make_reference([("chr1", 100_000)], seed=1, n_families=2)
    text, startpos, sizes, descr = sequence.text_from_chromosomes(ref.names, ref.seqs, rcref=True)
    ctx = api.Context.from_text(text, startpos, sizes, descr, keep_isa=False)
    batch = synth.make_reads_fast(ref.concat(), args.pairs, read_len=150, seed=1000)
    if args.fastq:
        t = fastq_texts(batch)
        pins = [api.PinnedArray(x.shape, np.uint8) for x in t]
        for p, x in zip(pins, t):
            p.array[...] = x
        call = lambda: ctx.text_upload(api.TEXT_FASTQ_PAIR, pins[0].array, pins[1].array, replace_n=True)   # noqa: E731
        nbytes = sum(x.size for x in t)
    else:
        t = synth.sam_text_fast(batch)
        pin = api.PinnedArray(t.shape, np.uint8)
        pin.array[...] = t
        call = lambda: ctx.text_upload(api.TEXT_SAM, pin.array)                                             # noqa: E731
        nbytes = t.size
    n, _ = call()                                                   # warm-up: allocations
    got = ctx.fetch_batch(0)
    want_seq = np.where(batch.seq == ord("N"), ord("Z"), batch.seq) if args.fastq else batch.seq     # N -> Z (fastqs_to_sam.cpp:69)
    exact = bool(n == batch.n and np.array_equal(got.seq, want_seq) and np.array_equal(got.qual, batch.qual)
                 and np.array_equal(got.names, batch.names) and np.array_equal(got.read_flag, api.read_flags_from_sam_flags(batch.flags)))
    ctx.ingest_ms(reset=True)
    l0 = ctx.launches
    t0 = time.perf_counter()
    for _ in range(args.calls):
        call()
    wall = (time.perf_counter() - t0) * 1e3 / args.calls
    dev = ctx.ingest_ms(reset=True) / args.calls
    out_bytes = batch.names.size + 2 * batch.seq.size + 18 * batch.n
    print(json.dumps({"input": "fastq pair" if args.fastq else "sam", "reads": int(n), "text_bytes": int(nbytes), "calls": args.calls,
                      "device_ms": dev, "wall_ms": wall, "launches_per_call": (ctx.launches - l0) // args.calls,
                      "reads_per_s_device": n / (dev / 1e3), "reads_per_s_wall": n / (wall / 1e3),
                      "alg_gbs": (nbytes + out_bytes) / (dev / 1e3) / 1e9, "h2d_gbs_wall": nbytes / (wall / 1e3) / 1e9,
                      "matches_generated_batch": exact}))
    ctx.close()


if __name__ == "__main__":
    main()
