#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] as measured lines on one B200 (measurement tooling, not product code).

    python profiles/sweeps.py --workload config2 > gpurun_out/sweeps.jsonl

configs[3]  bin-resolution sweep: the same mapped reads binned into 50 k / 100 k / 500 k bins (the reference ships only
            sample_bins/50000; the finer sets split every 50 k bin into 2 / 10 equal parts) with the
            mappability filter on -- time of smash_tail_finish (dedupe + filters + histogram) and the counts' checksum;
configs[4]  read-length / fragment-density / min-length sweep: 100 / 150 / 250 bp reads, 3-8 fragments per read,
            `-l` 16 / 20 / 24 -- device ms per 1 M-read step (CUDA events inside the library), stage split, records per read.
MEM mode (`-maxmatch`, longSA::MEM) is measured on config1 (`--workload config1 --mem`): at hg19 scale the ISA it needs
does not leave room for the batch buffers next to the 8-byte index on one 180 GB GPU.
One JSON line per point.  Every point is a fresh context (the index is rebuilt on the GPU, ~10 s at hg19 scale) because
the seed/ext tables depend on the minimum length.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split_bins(starts, parts, total):
    """Every bin [s_i, s_{i+1}) cut into `parts` equal parts (integer division, remainders to the last part)."""
    ends = np.append(starts[1:], total)
    out = []
    for p in range(parts):
        out.append(starts + (ends - starts) * p // parts)
    return np.stack(out, axis=1).reshape(-1).astype(np.int64)        # bin-major, part-minor (monotone: bins do not overlap)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch-reads", type=int, default=1_000_000)
    ap.add_argument("--min-lens", default="16,20,24")
    ap.add_argument("--read-lens", default="100,150,250")
    ap.add_argument("--bins", default="1,2,10")
    ap.add_argument("--mem", action="store_true", help="also time -maxmatch (MEM mode) at min length 20")
    args = ap.parse_args()
    import bench
    from smash_paper_b200 import api, sequence, synth
    wl = bench.WORKLOADS[args.workload]
    ref = synth.make_reference(wl["chroms"], seed=1, n_families=wl["families"])
    text, startpos, sizes, descr = sequence.text_from_chromosomes(ref.names, ref.seqs, rcref=True)
    genome = ref.concat()
    workdir = "/tmp/smash_sweeps"
    os.makedirs(workdir, exist_ok=True)
    starts50 = bench.make_bins(wl, ref, workdir)
    offs = ref.offsets()
    B = args.batch_reads
    read_lens = [int(x) for x in args.read_lens.split(",")]
    batches = {q: [synth.make_reads_fast(genome, B // 2, read_len=q, seed=1000, first_pair=i * (B // 2)) for i in range(2)] for q in read_lens}
    del genome
    points = [(api.MODE_MAM, int(L)) for L in args.min_lens.split(",")]
    if args.mem:
        points.append((api.MODE_MEM, 20))
    for mode, L in points:
        t0 = time.time()
        mem = mode == api.MODE_MEM
        ctx = api.Context.from_text(text, startpos, sizes, descr, keep_isa=True, chunk_cap=wl["chunk_cap"], mode=mode,
                                    min_len=L, nomap=True, tag_mappability=not mem)
        if not mem:
            ctx.build_mappability_device(ref.total)
            ctx.drop_isa()
        t_index = time.time() - t0
        want = api.WANT_SAM if mem else api.WANT_SAM | api.WANT_TAIL
        if not mem:
            ctx.tail_configure(starts50, ref.names, offs)
            ctx.tail_reserve((B // 2) * (args.steps + 3), 8 * B * (args.steps + 3))
        for q in read_lens:
            if mem and q != wl["read_len"]:
                continue
            bs = batches[q]
            for i in range(2):
                ctx.upload(bs[i % 2]); ctx.map_resident(want)
            if not mem:
                ctx.tail_reset()
            ctx.stage_ms(reset=True)
            ms, nrec, sam = 0.0, 0, 0
            for i in range(args.steps):
                ctx.upload(bs[i % 2])
                r = ctx.map_resident(want)
                ms += r.gpu_ms; nrec += r.n_records; sam += r.sam_bytes
            st = ctx.stage_ms(reset=True)
            line = {"sweep": "mem_mode" if mem else "read_len_min_len", "workload": args.workload, "mode": "MEM" if mem else "MAM",
                    "min_len": L, "read_len": q, "fragments_per_read": "3-8", "reads_per_step": B, "steps": args.steps,
                    "ms_per_step": ms / args.steps, "reads_per_s_device": B / (ms / args.steps / 1e3),
                    "records_per_read": nrec / (B * args.steps), "sam_bytes_per_read": sam / (B * args.steps),
                    "stage_ms": {k: round(v / args.steps, 4) for k, v in st.items()}, "index_s": round(t_index, 1)}
            print(json.dumps(line), flush=True)
            if not mem and L == wl["min_len"] and q == wl["read_len"]:
                # configs[3]: the pairs just appended, binned at three resolutions
                for parts in [int(x) for x in args.bins.split(",")]:
                    starts = starts50 if parts == 1 else split_bins(starts50, parts, int(ref.total))
                    ctx.tail_configure(starts, ref.names, offs)
                    ctx.tail_reserve((B // 2) * (args.steps + 3), 8 * B * (args.steps + 3))
                    for rep in range(2):                    # first pass warms the finish kernels at this bin count
                        ctx.tail_reset()
                        for i in range(args.steps):
                            ctx.upload(bs[i % 2]); ctx.map_resident(want)
                        if rep == 0:
                            ctx.tail_finish()
                    t1 = time.perf_counter()
                    counts, stats = ctx.tail_finish()
                    fin = (time.perf_counter() - t1) * 1e3
                    counts = np.asarray(counts)
                    if parts == 1:
                        counts50 = counts.copy()
                    # a finer set refines the 50 k set: its counts summed per parent bin are the 50 k counts
                    folds = bool(np.array_equal(counts.reshape(-1, parts).sum(axis=1), counts50))
                    print(json.dumps({"sweep": "bins", "workload": args.workload, "n_bins": int(len(starts)), "reads": B * args.steps,
                                      "tail_finish_ms": fin, "counted": int(counts.sum()), "counts_fold_to_50k": folds,
                                      "stats": stats}), flush=True)
                ctx.tail_configure(starts50, ref.names, offs)
                ctx.tail_reserve((B // 2) * (args.steps + 3), 8 * B * (args.steps + 3))
            if not mem:
                ctx.tail_reset()
        ctx.close()


if __name__ == "__main__":
    main()
