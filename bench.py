#!/usr/bin/env python
"""bench.py -- reads/s mapped (MAM search + SAM records) and binned (smashMEM filter + varbin) on
N B200s, next to the reference's CPU path on the box's own host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config1|config2] [--impl ours|reference]
                  [--scaling strong --total-reads 100000000]     (configs[2]: a fixed job sharded over the GPUs)

One "step" = one batch of synthetic reads through the whole hot path.  Prints ONE JSON line
(rank 0).  `value` = whole-job reads/s with the batch resident in HBM (CUDA events);
`e2e` = the same through smash_submit/smash_wait with pinned HOST buffers, H2D and D2H copies of
every step inside the timed region.  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from smash_paper_b200 import sequence, synth  # noqa: E402

METRIC = "reads_per_s_mapped_and_binned"
UNIT = "reads/s"

WORKLOADS = {
    # BASELINE.json configs[0]: 5 x 10 Mb random reference with planted repeats, 150 bp reads, l=20, 50 kb bins
    "config1": dict(chroms=[(f"chr{i + 1}", 10_000_000) for i in range(5)], read_len=150, min_len=20,
                    bin_width=50_000, families=200, chunk_cap=0,
                    # bytes the REFERENCE algorithm touches per read (w=4)
                    ref_alg_bytes_per_read=7800.0),
    # configs[1]: hg19-shaped 24 chromosomes (3.1 Gb), 8-byte index
    "config2": dict(chroms=synth.HG19_SIZES, read_len=150, min_len=20, bin_width=None, families=2000,
                    chunk_cap=400_000_000, ref_alg_bytes_per_read=14100.0),
    # tiny, for CI
    "tiny": dict(chroms=[("chr1", 300_000), ("chr2", 200_000)], read_len=150, min_len=20, bin_width=5000,
                 families=10, chunk_cap=0, ref_alg_bytes_per_read=7800.0),
}


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        self.gpu = gpu
        self.proc = None
        self.path = None

    def start(self):
        self.path = tempfile.mktemp(suffix=".csv")
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pw.append(0.0)
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            # "under load": the samples of the upper half by power draw (the sampler also sees the untimed uploads)
            cut = float(np.median(pw)) if pw else 0.0
            loaded = [c for c, p_ in zip(sm, pw) if p_ >= cut] or sm
            out.update(sm_mhz=float(np.median(loaded)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def make_reference(wl, seed=1):
    t0 = time.time()
    ref = synth.make_reference(wl["chroms"], seed=seed, n_families=wl["families"])
    log(f"reference: {len(ref.names)} chromosomes, {ref.total} bp in {time.time() - t0:.1f}s")
    return ref


def make_bins(wl, ref, workdir):
    """bins: start_abspos column.  config2 uses the reference's own sample_bins/50000/bins.txt
    (committed copy under tests/golden/, hg19 offsets); the others fixed-width bins."""
    if wl["bin_width"] is None:
        path = os.path.join(ROOT, "tests", "golden", "sample_bins_50000_bins.txt")
        starts = np.array([int(line.split("\t")[2]) for line in open(path)], dtype=np.int64)
        return starts
    path = os.path.join(workdir, "bins.txt")
    synth.write_fixed_bins(ref, path, wl["bin_width"])
    return np.array([int(line.split("\t")[2]) for line in open(path)], dtype=np.int64)


def pinned_batch(api, batch):
    """Copy a synth.ReadBatch into cudaHostAlloc memory."""
    out, keep = {}, []
    for k in ("names", "name_off", "seq", "qual", "seq_off", "flags", "opt", "opt_off"):
        a = getattr(batch, k)
        p = api.PinnedArray(a.shape, a.dtype)
        p.array[...] = a
        out[k] = p.array
        keep.append(p)
    b = synth.ReadBatch(**out)
    rf = api.PinnedArray(batch.flags.shape, np.uint16)        # reader-side flag word (query.h:56-64), 2 B/read
    rf.array[...] = api.read_flags_from_sam_flags(batch.flags)
    b.read_flag = rf.array
    keep.append(rf)
    b._pinned = keep
    return b


# ------------------------------------------------------------------------------------------ ours

def run_ours(args, wl, rank, world):
    import torch
    from smash_paper_b200 import api
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    workdir = tempfile.mkdtemp(prefix="smash_bench_")
    ref = make_reference(wl)
    names = ref.names
    t0 = time.time()
    text, startpos, sizes, descr = sequence.text_from_chromosomes(names, ref.seqs, rcref=True)
    log(f"text: N={len(text)} in {time.time() - t0:.1f}s")
    t0 = time.time()
    ctx = api.Context.from_text(text, startpos, sizes, descr, keep_isa=True, chunk_cap=wl["chunk_cap"], device=dev,
                                min_len=wl["min_len"], nomap=True, tag_mappability=True)
    t_index = time.time() - t0
    n_text = int(len(text))
    del text                                                             # host copy no longer needed (6.2 GB at hg19 scale)
    t0 = time.time()
    ctx.build_mappability_device(ref.total)
    t_map = time.time() - t0
    log(f"index built on GPU in {t_index:.1f}s, map.bin in {t_map:.1f}s, {ctx.index_bytes / 1e9:.2f} GB in HBM")
    ref_files = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            ref_files = prepare_reference_files(ref, ctx, workdir)       # needs the ISA, so before it is dropped
        except Exception as e:  # noqa: BLE001
            ref_files = {"error": str(e)[:300]}
    ctx.drop_isa()                                                       # 49.5 GB at hg19 scale
    log(f"{ctx.index_bytes / 1e9:.2f} GB of index in HBM for the run")
    starts = make_bins(wl, ref, workdir)
    offs = ref.offsets()
    ctx.tail_configure(starts, names, offs)
    # capacity hint: this rank's pairs/hits, and room in the dedupe table for the other ranks' keys
    ctx.tail_reserve((args.batch_reads // 2) * (args.steps + 1) * world, 8 * args.batch_reads * (args.steps + 1))
    genome = ref.concat()

    B = args.batch_reads
    n_batches = args.warmup + args.steps
    pairs_per_batch = B // 2
    batches = []
    t0 = time.time()
    for i in range(n_batches):
        first = (rank * n_batches + i) * pairs_per_batch
        b = synth.make_reads_fast(genome, pairs_per_batch, read_len=wl["read_len"], seed=1000, first_pair=first)
        batches.append(pinned_batch(api, b))
    log(f"{n_batches} batches x {B} reads generated in {time.time() - t0:.1f}s")
    del genome
    if world > 1:
        ref.seqs = []                                                    # keep host RAM per rank small at 8 ranks
    want = api.WANT_SAM | api.WANT_TAIL

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident: inputs in HBM before the timed region of every step
    counts_t = torch.zeros(len(starts), dtype=torch.int64, device=f"cuda:{dev}")
    from smash_paper_b200 import multigpu
    backend = multigpu.ContextBackend(ctx, rank * args.steps * pairs_per_batch, torch.device("cuda", dev))
    torch_tail = bool(os.environ.get("SMASH_TORCH_TAIL"))       # A/B: round 1's exchange in torch.distributed (multigpu.py)
    if dist and not torch_tail:
        # NCCL behind the C ABI (csrc/comm.cu): rank 0's unique id goes round through the job's rendezvous store
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_rank(rank, world, uid[0])

    def finish():
        """smashMEM dedupe + varbin over everything accumulated; across ranks smash_bins_finish: hash-partitioned
        exchange of the dupe fingerprints (two all-to-alls, resolved on the device), all-gather of the shard edges and
        ONE ncclAllReduce of the per-bin counts -- exact, not per-shard."""
        if dist and torch_tail:
            return multigpu.sharded_tail_finish(backend, dist, rank, world)
        if dist:
            _, st_ = ctx.bins_finish(ordinal_base=rank << 40, counts_device_ptr=counts_t.data_ptr())
            return counts_t, st_
        c, st_ = ctx.tail_finish(counts_t.data_ptr())
        return counts_t, st_

    sampler = ClockSampler(dev)
    sampler.start()                                      # samples from the warm-up to the end of the e2e region
    for i in range(args.warmup):
        ctx.upload(batches[i]); ctx.map_resident(want)
    # untimed sizing pass: the tail must have seen as many pairs as the timed region will give it, so that no
    # buffer (library, torch allocator, NCCL) grows inside the timed tail_finish -- a growing rank makes the
    # others wait in the exchange
    for i in range(args.warmup, max(args.warmup, args.steps)):
        ctx.upload(batches[i % len(batches)]); ctx.map_resident(want)
    finish()                                             # warm-up of the tail kernels and the collectives too
    ctx.tail_reset(); ctx.stage_ms(reset=True)
    barrier()
    launches0 = ctx.launches
    dev_ms = 0.0
    sam_bytes = 0
    stats_nrec = 0
    for i in range(args.steps):
        ctx.upload(batches[args.warmup + i])            # untimed: H2D, then the step runs on resident data
        r = ctx.map_resident(want)
        dev_ms += r.gpu_ms
        sam_bytes += r.sam_bytes
        stats_nrec += r.n_records
    # tail_finish runs on the library's own stream and returns synchronised; the allreduce runs on
    # torch's stream: time both on the host between two full synchronisations
    barrier()                                            # ranks reach the tail together (their untimed uploads differ)
    t_f = time.perf_counter()
    counts_g, stats = finish()                           # includes the one all_reduce of the per-bin counts
    torch.cuda.synchronize()
    finish_ms = (time.perf_counter() - t_f) * 1e3
    dev_ms += finish_ms
    stage = ctx.stage_ms()
    launches = ctx.launches - launches0
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{dev}")
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t_ms.item())
    total_reads = B * args.steps * world
    value = total_reads / (dev_ms_max / 1e3)

    # ---------------- end to end: pinned host batches in, SAM bytes + counts back on the host
    host_cores = len(os.sched_getaffinity(0))
    host_threads = max(1, min(16, host_cores // max(world, 1)))

    def e2e_loop(full_sam_text):
        """smash_submit / smash_wait over both slots.  Default transport: the text the GPU computes + 32 B per record come
        back and `host_threads` library threads rebuild the byte-identical SAM lines in host memory from the submitted
        batch; full_sam_text: the whole SAM text crosses PCIe (round 1's path, kept as the A/B line)."""
        ctx.set_transport(full_sam_text=full_sam_text, host_threads=host_threads)
        ctx.tail_reset()
        for i in range(max(api.N_SLOTS, min(2 * api.N_SLOTS, args.warmup))):   # warm every slot (buffers, pinned results, threads)
            ctx.submit(i % api.N_SLOTS, batches[i % len(batches)], want=want)
            ctx.wait(i % api.N_SLOTS, copy=False)
        ctx.tail_reset()
        ctx.io_bytes(reset=True)
        barrier()
        t0 = time.perf_counter()
        sam_out = 0
        for i in range(args.steps):
            slot = i % api.N_SLOTS
            if i >= api.N_SLOTS:
                sam_out += int(ctx.wait(slot, copy=False).sam_bytes)
            ctx.submit(slot, batches[args.warmup + i], want=want, first_pair=i * pairs_per_batch)
        for i in range(max(0, args.steps - api.N_SLOTS), args.steps):
            sam_out += int(ctx.wait(i % api.N_SLOTS, copy=False).sam_bytes)
        counts_e, stats_e = finish()
        host_counts_e = counts_e.cpu()
        barrier()
        t_e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{dev}")
        if dist:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        h2d_b, d2h_b = ctx.io_bytes(reset=True)
        return total_reads / float(t_e.item()), h2d_b, d2h_b + host_counts_e.numel() * 8, sam_out, host_counts_e

    e2e_full_value, h2d_full, d2h_full, _, host_counts_full = e2e_loop(True)
    e2e_value, h2d, d2h, sam_out_e2e, host_counts = e2e_loop(False)
    counts_equal_full = bool(torch.equal(host_counts_full, host_counts))

    # ---------------- the same loop for a caller that archives no mapout: reads in, bin counts out (WANT_TAIL only).
    # The production pipeline's product is the bin counts (binning.sh); the SAM text is an intermediate file.
    ctx.tail_reset()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        slot = i % api.N_SLOTS
        if i >= api.N_SLOTS:
            ctx.wait(slot, copy=False)
        ctx.submit(slot, batches[args.warmup + i], want=api.WANT_TAIL, first_pair=i * pairs_per_batch)
    for i in range(max(0, args.steps - api.N_SLOTS), args.steps):
        ctx.wait(i % api.N_SLOTS, copy=False)
    counts_g3, stats3 = finish()
    host_counts3 = counts_g3.cpu()
    barrier()
    t_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{dev}")
    if dist:
        dist.all_reduce(t_t, op=dist.ReduceOp.MAX)
    e2e_tail_value = total_reads / float(t_t.item())
    tail_counts_equal = bool(torch.equal(host_counts3, host_counts))
    clocks = sampler.stop()
    dma = dma_ceiling(torch, dev, dist)

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # roofline of the DOMINANT kernel of the step; algorithmic (element-granular) bytes per launch as
        # defined in DESIGN.md §4 for each kernel
        n_rec_per_read = stats_nrec / max(B * args.steps, 1)
        sam_per_read = sam_bytes / max(B * args.steps, 1)
        split = stage.get("verify", 0.0) > 0.0                  # split search: k_mam_search parks, k_mam_verify extends
        alg = kernel_alg_bytes(wl, n_text, n_rec_per_read, sam_per_read, split)
        # the "search" stage timer spans k_mam_seed (lanes = anchors; nearly all of the time) + k_mam_search (exact paths of
        # the few flagged reads); without the split search it is k_mam_search alone
        SEARCH_KEY = "k_mam_seed+k_mam_search" if split else "k_mam_search"
        alg[SEARCH_KEY] = alg["k_mam_search"]
        per_kernel = {}
        for kname, skey in ((SEARCH_KEY, "search"), ("k_mam_verify", "verify"), ("k_rec_build+k_rec_xe", "records"),
                            ("k_sizes", "sizes_scan"), ("k_emit_text", "emit_text"), ("k_emit_copy", "emit_copy")):
            if kname == "k_mam_verify" and not split:
                continue
            ms = stage[skey] / args.steps
            gbs = B * alg[kname] / (ms / 1e3) / 1e9 if ms > 0 else 0.0
            per_kernel[kname] = {"ms": ms, "alg_bytes_per_read": alg[kname], "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
        # DRAM traffic of that kernel from the committed ncu --set full capture of this workload (1 M reads per launch)
        traffic, traffic_src = None, None
        try:
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            from kernel_hash import kernel_source_hash
            tj = json.load(open(os.path.join(ROOT, "profiles", f"r02_dram_traffic_{args.workload}.json")))
            parts = [k for k in dom.split("+") if k in tj]
            if tj.get("_kernel_sources_sha256") != kernel_source_hash():
                traffic_src = "stale: the kernel sources changed after profiles/r02_dram_traffic_*.json was captured (field dropped)"
            elif parts:
                traffic = sum(tj[k]["dram_read_bytes"] + tj[k]["dram_write_bytes"] for k in parts) * (B / 1e6)
                traffic_src = (f"profiles/r02_ncu_full_{args.workload}_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum of {' + '.join(parts)}, "
                               f"scaled to {B} reads/launch; capture taken from these kernel sources, sha256 checked)")
        except (OSError, ValueError):
            pass
        search_ms = per_kernel[SEARCH_KEY]["ms"] + (per_kernel["k_mam_verify"]["ms"] if split else 0.0)
        achieved = per_kernel[dom]["achieved_gbs"]
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload_desc, "reads_per_step_per_gpu": B, "read_len": wl["read_len"],
                       "min_len": wl["min_len"], "text_len": int(n_text), "n_bins": int(len(starts)),
                       "total_reads": int(total_reads), "index": "built on GPU, replicated per GPU", "sharding": f"reads x{world} (contiguous pair ranges, index replicated); tail exact across shards inside smash_bins_finish (C ABI, NCCL): partitioned exchange of dupe fingerprints, all-gather of shard edges, 1 ncclAllReduce of bin counts",
                       "l2": "inputs (index touches, 1.7 KB/read SAM) far larger than L2; distinct batch per step",
                       "timing": "sum of per-step CUDA-event durations with the batch resident + tail_finish/allreduce; max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(args.steps, 1),
                    "d2h_bytes_per_step": d2h // max(args.steps, 1), "want": "SAM text + bin counts (SMASH_WANT_SAM | SMASH_WANT_TAIL)",
                    "transport": "compact: head/tags/L-R text + 32 B per record over PCIe, SAM lines rebuilt in host memory by the library's host threads (byte-identical, tests/test_transport.py)",
                    "host_threads": host_threads, "host_cores": host_cores,
                    "sam_bytes_in_host_memory_per_step": sam_out_e2e // max(args.steps, 1),
                    "d2h_gbs_per_gpu": (d2h / max(args.steps, 1)) * (e2e_value / world / B) / 1e9,
                    "frac_of_d2h_ceiling": ((d2h / max(args.steps, 1)) * (e2e_value / world / B) / 1e9) / dma["d2h_gbs_concurrent_per_gpu"]
                    if dma.get("d2h_gbs_concurrent_per_gpu") else None},
            "e2e_full_sam_text": {"value": e2e_full_value, "unit": UNIT, "h2d_bytes_per_step": h2d_full // max(args.steps, 1),
                                  "d2h_bytes_per_step": d2h_full // max(args.steps, 1), "counts_equal_compact_run": counts_equal_full,
                                  "transport": "whole SAM text over PCIe (smash_ctx_set_transport(full_sam_text=1)): round 1's path, A/B line",
                                  "d2h_gbs_per_gpu": (d2h_full / max(args.steps, 1)) * (e2e_full_value / world / B) / 1e9},
            "e2e_tail_only": {"value": e2e_tail_value, "unit": UNIT, "h2d_bytes_per_step": h2d // max(args.steps, 1),
                              "d2h_bytes_per_step": int(host_counts.numel() * 8 // max(args.steps, 1)),
                              "want": "bin counts only (SMASH_WANT_TAIL): no mapout text leaves the GPU", "counts_equal_sam_run": tail_counts_equal},
            "dma_ceiling": dma,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": dom, "kernel_ms": per_kernel[dom]["ms"], "alg_bytes_per_read": per_kernel[dom]["alg_bytes_per_read"],
                         "kernels": per_kernel,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "ref_alg_bytes_per_read": wl["ref_alg_bytes_per_read"],
                         "search_on_ref_alg_bytes_gbs": B * wl["ref_alg_bytes_per_read"] / (search_ms / 1e3) / 1e9 if search_ms > 0 else None,
                         "split_search": bool(split),
                         # the dominant kernel fetches random 128-byte lines: its memory ceiling is the random-line rate of
                         # the GPU at the kernel's footprint, measured by profiles/microbench/random_gather.cu (committed
                         # numbers, not re-measured here), next to which the ncu DRAM traffic of the kernel is placed
                         "random_line_ceiling": {"gbs": 4640.0, "glines_per_s": 36.3, "valid_up_to_footprint_gb": 68,
                                                 "source": "profiles/r02_random_gather.md",
                                                 "traffic_frac": (traffic / (per_kernel[dom]["ms"] / 1e3) / 1e9 / 4640.0) if traffic else None}},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "sam_bytes_per_read": sam_bytes / max(B * args.steps, 1),
            "tail": stats, "tail_finish_ms": finish_ms, "index_build_s": t_index, "mappability_build_s": t_map,
            "index_hbm_gb": ctx.index_bytes / 1e9,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                if "error" in ref_files:
                    raise RuntimeError(ref_files["error"])
                out["cpu_baseline"] = cpu_baseline(args, wl, ref, ref_files, sample_pairs=args.cpu_sample_pairs, steps=1, check_ctx=ctx)
                out["parity_check"] = out["cpu_baseline"].pop("parity_check")
                rc = out["cpu_baseline"].pop("ref_counters", None)
                if rc and "alg_bytes_per_read" in rc:
                    # measured on this workload by the counting build of the reference: replaces the survey's estimate
                    rl = out["roofline"]
                    rl["ref_alg_bytes_per_read_estimate"] = rl["ref_alg_bytes_per_read"]
                    rl["ref_alg_bytes_per_read"] = rc["alg_bytes_per_read"]
                    rl["ref_alg_bytes_source"] = "measured: oracle/_ref counting build on a sample of this workload (ref_counters)"
                    if rl.get("search_on_ref_alg_bytes_gbs"):
                        rl["search_on_ref_alg_bytes_gbs"] *= rc["alg_bytes_per_read"] / rl["ref_alg_bytes_per_read_estimate"]
                    rl["step_on_ref_alg_bytes_frac"] = out["value"] / world * rc["alg_bytes_per_read"] / 1e9 / rl["peak"] if rl.get("peak") else None
                out["ref_counters"] = rc
            except Exception as e:  # noqa: BLE001 -- the baseline must never break the bench line
                out["cpu_baseline"] = {"error": str(e)[:300]}
    if out is not None and world == 1:
        # input stage, measured last and on its own so that it can never disturb the numbers above
        try:
            out["ingest"] = measure_ingest(api, ctx, batches[args.warmup], peak if rank == 0 else 6650.0, steps=min(args.steps, 5))
        except Exception as e:  # noqa: BLE001
            out["ingest"] = {"error": str(e)[:300]}
    ctx.close()
    shutil.rmtree(workdir, ignore_errors=True)
    if dist:
        dist.destroy_process_group()
    return out


def dma_ceiling(torch, dev, dist, mb=256, reps=6):
    """What the box's host<->device DMA can do for THIS rank while every rank does the same: pinned cudaMemcpyAsync of
    256 MB buffers, D2H alone, H2D alone and both directions at once (CUDA events, all ranks started together)."""
    n = mb << 20
    try:
        hp_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); hp_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d_in = torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}"); d_out = torch.zeros(n, dtype=torch.uint8, device=f"cuda:{dev}")
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def run(h2d, d2h):
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_event(e0); s2.wait_event(e0)
            with torch.cuda.stream(s1):
                for _ in range(reps if h2d else 0):
                    d_in.copy_(hp_in, non_blocking=True)
                e1.record()
            with torch.cuda.stream(s2):
                for _ in range(reps if d2h else 0):
                    hp_out.copy_(d_out, non_blocking=True)
                e2.record()
            torch.cuda.synchronize()
            return (reps * n / (e0.elapsed_time(e1) / 1e3) / 1e9 if h2d else None, reps * n / (e0.elapsed_time(e2) / 1e3) / 1e9 if d2h else None)

        run(True, True)
        h_alone = run(True, False)[0]
        d_alone = run(False, True)[1]
        h_both, d_both = run(True, True)
        vals = torch.tensor([h_alone, d_alone, h_both, d_both], dtype=torch.float64, device=f"cuda:{dev}")
        if dist:
            dist.all_reduce(vals, op=dist.ReduceOp.MIN)          # the slowest rank's view (ranks run concurrently)
        h_alone, d_alone, h_both, d_both = [float(x) for x in vals.tolist()]
        return {"what": f"pinned cudaMemcpyAsync, {mb} MB x {reps}, every rank at once; min over ranks", "h2d_gbs_alone_per_gpu": h_alone,
                "d2h_gbs_alone_per_gpu": d_alone, "h2d_gbs_concurrent_per_gpu": h_both, "d2h_gbs_concurrent_per_gpu": d_both}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def measure_ingest(api, ctx, batch, peak, steps):
    """Raw SAM text (what fastqs_to_sam pipes into `mummer -samin`, smash_mapping.sh:19) of one batch, in pinned host
    memory -> packed batch in HBM, parsed by ingest.cu.  wall = H2D copy + parse (host clock around the call);
    device = CUDA events from the end of the copy to the end of k_ing_copy.  Algorithmic bytes per read: the text
    once in, the batch (names, SEQ, QUAL, two 8-byte offsets, 2-byte flag) once out."""
    text = synth.sam_text_fast(batch)
    pin = api.PinnedArray(text.shape, np.uint8)
    pin.array[...] = text
    n = batch.n
    for _ in range(2):
        got, cons = ctx.text_upload(api.TEXT_SAM, pin.array)
    parsed = ctx.fetch_batch(0)
    exact = bool(got == n and cons[0] == text.size and np.array_equal(parsed.names, batch.names) and np.array_equal(parsed.seq, batch.seq)
                 and np.array_equal(parsed.qual, batch.qual) and np.array_equal(parsed.seq_off, batch.seq_off)
                 and np.array_equal(parsed.read_flag, api.read_flags_from_sam_flags(batch.flags)))
    ctx.ingest_ms(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.text_upload(api.TEXT_SAM, pin.array)
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    dev_ms = ctx.ingest_ms(reset=True) / steps
    # the whole drop-in path of `mummer -samin` from text: SAM text in pinned host memory -> SAM records in pinned host
    # memory, both slots in flight (the next chunk is parsed while the previous one is searched / downloaded)
    for i in range(api.N_SLOTS):
        ctx.submit_text(i, api.TEXT_SAM, pin.array, want=api.WANT_SAM)
    for i in range(api.N_SLOTS):
        ctx.wait(i, copy=False)
    t0 = time.perf_counter()
    sam_out = 0
    for i in range(steps):
        slot = i % api.N_SLOTS
        if i >= api.N_SLOTS:
            sam_out += int(ctx.wait(slot, copy=False).sam_bytes)
        ctx.submit_text(slot, api.TEXT_SAM, pin.array, want=api.WANT_SAM, first_pair=i * (n // 2))
    for i in range(max(0, steps - api.N_SLOTS), steps):
        sam_out += int(ctx.wait(i % api.N_SLOTS, copy=False).sam_bytes)
    text_to_sam_s = (time.perf_counter() - t0) / steps
    ctx.ingest_ms(reset=True)
    alg = text.size / n + (batch.names.size + 2 * batch.seq.size) / n + 18
    gbs = n * alg / (dev_ms / 1e3) / 1e9 if dev_ms > 0 else 0.0
    pin.free()
    return {"what": "SAM text -> packed batch on the device (smash_text_upload), 1 batch", "reads_per_step": int(n), "steps": steps,
            "text_bytes_per_read": text.size / n, "wall_ms_per_step": wall_ms, "device_ms_per_step": dev_ms,
            "reads_per_s_wall": n / (wall_ms / 1e3), "reads_per_s_device": n / (dev_ms / 1e3) if dev_ms > 0 else None,
            "alg_bytes_per_read": alg, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak if peak else None,
            "matches_generated_batch": exact,
            "text_to_sam": {"what": "smash_submit_text/smash_wait, 2 slots: SAM text (host) -> SAM records (host), no tail", "reads_per_s": n / text_to_sam_s,
                            "ms_per_step": text_to_sam_s * 1e3, "h2d_bytes_per_step": int(text.size), "d2h_bytes_per_step": sam_out // max(steps, 1)},
            "note": "device time spans 3 host round trips (line count, totals) and 9 kernels; wall adds the H2D copy of the text"}


def kernel_alg_bytes(wl, N, n_rec, sam_bytes, split=False):
    """Element-granular bytes each kernel must touch per READ.  n_rec = records/read and
    sam_bytes = SAM bytes/read are measured in the run; the candidate count is the workload's expectation."""
    import math
    w = 4 if N < 0xFFFFFFFF - 100000 else 8
    L, q = wl["min_len"], wl["read_len"]
    k = min(16, max(4, math.ceil(math.log(N) / math.log(4.0)) + 1), L)
    s = L - k + 1
    anchors = (q - L + s - 1) // s + 1
    seed_w = 4 if N < 0xFFFFFFFF else 8
    frag = q / 5.5                                                  # 3..8 fragments per read
    true_cand = max(0.0, frag - k + 1) / frag                      # anchor k-mer inside one fragment -> its locus
    cand = anchors * (N / 4.0 ** k + true_cand)
    name = 10
    surv = anchors * true_cand                                      # candidates left after the 4+4 filter (chance hits: ~2 %)
    if split:
        # read in + lower-cased copy out; per anchor the bucket bounds -- two entries of the flat seed table, or, in the blocked
        # table of the 8-byte case, the 24-byte block header the kernel loads (40-bit rank + 16 bucket sizes) --;
        # per bucket entry its 4-byte ext code; parked candidates out
        blocked = seed_w == 8 and k == 16
        search = q + q + anchors * (24 if blocked else 2 * seed_w) + cand * 4 + surv * 8
        # parked candidates in, their SA entry and the 8-byte text window left of the seed, the read once; per fragment
        # (the candidates that own a diagonal) its text span and one U byte; matches out
        verify = surv * (8 + w + 8) + q + (q / frag) * (frag + 1) + 16 * n_rec
    else:
        # read + per anchor two seed entries + per candidate (SA entry, two 16 B text windows, U byte) + matches out
        search = q + anchors * 2 * seed_w + cand * (w + 32 + 1) + 16 * n_rec
        verify = 0.0
    return {
        "k_mam_search": search,
        "k_mam_verify": verify,
        # matches in, Rec(40)+Item(4) out, per record the read and the text diagonal (XE) and 2 map bytes
        "k_rec_build+k_rec_xe": 16 * n_rec + 44 * n_rec + 2 * q * n_rec + 2 * n_rec + 16,
        # per record: its Rec + neighbours' Rec/Item for the cc/CC tags, 4 B out
        "k_sizes": n_rec * (3 * 40 + 12 + 4 + 16),
        # variable text out (everything except name/SEQ/QUAL); in: Rec/Item of the record and its neighbours, offsets
        "k_emit_text": (sam_bytes - n_rec * (name + 2 * q + 1)) + n_rec * (3 * 40 + 12 + 8 + 16),
        # bulk bytes out (name + SEQ + tab + QUAL per record); in: the read once (name + SEQ + QUAL) + Rec + offset per record
        "k_emit_copy": n_rec * (name + 2 * q + 1) + (name + 2 * q) + n_rec * (40 + 8),
    }


# ------------------------------------------------------------------------------------- reference

cleanup = []


def prepare_reference_files(ref, ctx, workdir):
    """FASTA + <fa>.bin/ index files for the reference binary.  ctx: a GPU context whose index is saved in
    the reference's file format (byte-identical files, tests/test_gpu_parity.py), or None -> the reference
    builds its own index (qsufsort).  Big indexes go to tmpfs (the box's disk is smaller than an hg19-scale index)."""
    from oracle import oracle as O
    if not O.have_reference():
        raise RuntimeError("oracle/_ref not built")
    cores = os.cpu_count() or 2
    N = 2 * ref.total + 2 * len(ref.names)
    w = 8 if N >= 0xFFFFFFFF - 100000 else 4
    need = N * (2 + 2 * w) + ref.total + (cores + 2) * 600_000_000      # index files + FASTA + mummer's 500 MB/thread arenas
    if need > 2_000_000_000:
        import psutil
        avail = psutil.virtual_memory().available
        if avail < need + 24_000_000_000 or not os.path.isdir("/dev/shm"):
            raise RuntimeError(f"not enough host RAM for the reference's index files: need {need / 1e9:.0f} GB, {avail / 1e9:.0f} GB available")
        workdir = tempfile.mkdtemp(prefix="smash_ref_", dir="/dev/shm")
        cleanup.append(workdir)
    fa = os.path.join(workdir, "ref.fa")
    t0 = time.time()
    synth.write_fasta(ref, fa)
    long_ints = w == 8
    if ctx is not None:
        ctx.save_index(fa, with_mappability=False)
        built = "gpu builder (files byte-identical to the reference's, see tests)"
    else:
        O.ref_build_index(fa, long_ints=long_ints, mappability=False)
        built = "reference's own qsufsort build"
    log(f"reference index files ready in {time.time() - t0:.1f}s ({built})")
    return {"fa": fa, "workdir": workdir, "long_ints": long_ints, "built": built}


def mapout_lines(workdir):
    """Record lines of every chunk file the reference wrote under <workdir>/mapout (headers dropped)."""
    import glob
    lines = []
    for fn in glob.glob(os.path.join(workdir, "mapout", "*.txt")):
        with open(fn, "rb") as f:
            lines += [ln for ln in f if not ln.startswith(b"@")]
    return lines


def parity_check(ctx, batch, ref_lines):
    """The reads the reference binary has just mapped, through the GPU path (smash_map_batch, host buffers in,
    SAM text out), compared line for line as sorted multisets (the reference's chunk membership and in-chunk order
    depend on thread scheduling).  Returns the JSON object of the bench line."""
    from smash_paper_b200 import api
    ctx.set_tag_mappability(False)                       # the reference's mapout has no L/R tags (a later stage adds them)
    try:
        res = ctx.map_batch(batch, want=api.WANT_SAM)
    finally:
        ctx.set_tag_mappability(True)
    ours = sorted(res.sam.splitlines(keepends=True))
    theirs = sorted(ref_lines)
    equal = ours == theirs
    out = {"what": "sorted SAM record lines: oracle/_ref binary vs smash_map_batch on the same reads, this workload's index",
           "reads": int(batch.n), "lines": len(theirs), "lines_gpu": len(ours), "equal": bool(equal)}
    if not equal:
        for a, b in zip(ours, theirs):
            if a != b:
                out["first_difference"] = {"gpu": a[:300].decode(errors="replace"), "reference": b[:300].decode(errors="replace")}
                break
    return out


def reference_counters(wl, genome, files, n_pairs):
    """the reference with event counters compiled in (oracle/_ref/mummer[-long]-counters, built
    by oracle/make_counters.py from a patched scratch copy) run on a small sample of THIS workload, to measure the
    element-granular bytes the reference algorithm touches per read:
      B_alg = (E+S)(w+1) + K + Lk*4w + q(1+rec) + 2q
    E edge probes, S binary-search steps (each = one SA entry of w bytes + one text byte), K LCP reads, Lk suffix links
    (4 SA/ISA entries each), rec records per read (the XE scan of q text bytes each), 2q = bases + qualities."""
    from oracle import oracle as O
    fa, workdir, long_ints = files["fa"], files["workdir"], files["long_ints"]
    exe = os.path.join(O.REF_BIN, "mummer-long-counters" if long_ints else "mummer-counters")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/*-counters not built (oracle/make_counters.py)")
    cores = os.cpu_count() or 2
    b = synth.make_reads_fast(genome, n_pairs, read_len=wl["read_len"], seed=1000, first_pair=700_000_000)
    sam = os.path.join(workdir, "counters.sam")
    synth.write_sam(b, sam)
    shutil.rmtree(os.path.join(workdir, "mapout"), ignore_errors=True)
    p = subprocess.run([exe, "-rcref", "-qthreads", str(max(2, cores)), "-nomap", "-samin", "-samout", fa, sam],
                       cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    os.unlink(sam)
    line = [ln for ln in p.stderr.decode(errors="replace").splitlines() if ln.startswith("# smash_counters")]
    if not line:
        raise RuntimeError("no counter line on stderr")
    c = {k: int(v) for k, v in (kv.split("=") for kv in line[-1].split()[2:])}
    n = c["reads"]
    if n != 2 * n_pairs:
        raise RuntimeError(f"counted {n} reads, sent {2 * n_pairs}")
    rec = len(mapout_lines(workdir)) / n
    w, q = (8 if long_ints else 4), wl["read_len"]
    E, S, K, Lk = c["edge"] / n, c["steps"] / n, c["lcp"] / n, c["links"] / n
    b_alg = (E + S) * (w + 1) + K + Lk * 4 * w + q * (1 + rec) + 2 * q
    return {"reads": n, "per_read": {"edge_probes": E, "bsearch_steps": S, "lcp_reads": K, "suffix_links": Lk,
                                     "traverse_calls": c["traverse"] / n, "matches": c["emit"] / n,
                                     "link_expansions_failed": c["linkfail"] / n, "records": rec},
            "w": w, "q": q, "alg_bytes_per_read": b_alg,
            "formula": "(E+S)(w+1) + K + Lk*4w + q(1+rec) + 2q  (SURVEY 8d), counters of SURVEY App. D measured on this workload"}


def cpu_baseline(args, wl, ref, files, sample_pairs, steps, check_ctx=None):
    """The UNMODIFIED reference (oracle/_ref/mummer[-long]) on the host cores, bounded sample per step.
    check_ctx: a GPU context on the same reference -> the first sample's mapout is kept and compared with the GPU
    path's records for the same reads (`parity_check`)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 2
    fa, workdir, long_ints, built = files["fa"], files["workdir"], files["long_ints"], files["built"]
    exe = os.path.join(O.REF_BIN, "mummer-long" if long_ints else "mummer")
    march = "x86-64-v3"
    try:                                                 # closest build to the reference's -march=native that this host can run
        cpu_flags = set(next(ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")).split())
        if {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= cpu_flags and os.path.exists(exe + "-v4"):
            exe, march = exe + "-v4", "x86-64-v4"
    except (OSError, StopIteration):
        pass
    empty = os.path.join(workdir, "empty.sam")
    open(empty, "w").close()

    def run(sam):
        shutil.rmtree(os.path.join(workdir, "mapout"), ignore_errors=True)
        t = time.perf_counter()
        subprocess.run([exe, "-rcref", "-qthreads", str(max(2, cores)), "-nomap", "-samin", "-samout", fa, sam],
                       cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return time.perf_counter() - t

    genome = ref.concat()
    parity = None
    # production flags (MAP_POPULATE): the zero-read run measures index mmap+populate and the per-thread
    # 500 MB arena initialisation, which is subtracted so that only mapping+SAM time remains
    run(empty)
    startup = run(empty)
    vals = []
    for s in range(steps):
        b = synth.make_reads_fast(genome, sample_pairs, read_len=wl["read_len"], seed=1000, first_pair=500_000_000 + s * sample_pairs)
        sam = os.path.join(workdir, f"sample{s}.sam")
        synth.write_sam(b, sam)
        wall = run(sam)
        if check_ctx is not None and s == 0:
            try:
                parity = parity_check(check_ctx, b, mapout_lines(workdir))
            except Exception as e:  # noqa: BLE001
                parity = {"equal": False, "error": str(e)[:300]}
            log(f"parity_check: {parity}")
        vals.append(2 * sample_pairs / max(wall - startup, 0.1 * wall))
        log(f"reference step {s}: {2 * sample_pairs} reads wall {wall:.2f}s startup {startup:.2f}s -> {vals[-1]:.0f} reads/s")
        os.unlink(sam)
    counters = None
    if check_ctx is not None:
        try:
            counters = reference_counters(wl, genome, files, n_pairs=min(sample_pairs, 25_000))
        except Exception as e:  # noqa: BLE001
            counters = {"error": str(e)[:300]}
        log(f"reference counters: {counters}")
    return {"value": float(np.mean(vals)), "unit": UNIT, "cores": cores, "kind": "reference", "ref_counters": counters,
            "sample": f"{2 * sample_pairs} reads/step x {steps} through oracle/_ref/{os.path.basename(exe)} -rcref -qthreads {max(2, cores)} "
                      f"-nomap -samin -samout; wall minus a zero-read run ({startup:.2f}s index mmap + buffer init); "
                      f"mapping+SAM only (mappability_tag/smashMEM/varbin stages not included); reference flags -Ofast -march={march}",
            "index": built, "per_step": vals, "parity_check": parity}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return None
    workdir = tempfile.mkdtemp(prefix="smash_ref_")
    ref = make_reference(wl)
    ctx = None
    N = 2 * ref.total + 2 * len(ref.names)
    if N > 400_000_000:
        # the reference's own index build would take hours here: use the GPU builder's files
        from smash_paper_b200 import api
        text, startpos, sizes, descr = sequence.text_from_chromosomes(ref.names, ref.seqs, rcref=True)
        ctx = api.Context.from_text(text, startpos, sizes, descr, keep_isa=True, chunk_cap=wl["chunk_cap"])
        del text
    try:
        files = prepare_reference_files(ref, ctx, workdir)
        if ctx is not None:
            ctx.close(); ctx = None
        cb = cpu_baseline(args, wl, ref, files, sample_pairs=args.cpu_sample_pairs, steps=args.warmup + args.steps)
    finally:
        if ctx is not None:
            ctx.close()
        shutil.rmtree(workdir, ignore_errors=True)
    vals = cb["per_step"][args.warmup:] or cb["per_step"]
    v = float(np.mean(vals))
    cb["value"] = v
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * 2 * args.cpu_sample_pairs / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload_desc, "read_len": wl["read_len"], "min_len": wl["min_len"],
                       "reads_per_step": 2 * args.cpu_sample_pairs},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    # stdout carries exactly ONE JSON line: everything else that might write to fd 1 (NCCL's version banner,
    # library chatter) is sent to stderr; the JSON goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SMASH_BENCH_WORKLOAD", "config2"), choices=sorted(WORKLOADS))
    ap.add_argument("--batch-reads", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample-pairs", type=int, default=150_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # BASELINE.json configs[2]: a FIXED number of reads sharded over the GPUs (strong scaling); --steps is then derived
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-reads", type=int, default=100_000_000)
    args = ap.parse_args()
    if args.scaling == "strong":
        world_ = int(os.environ.get("WORLD_SIZE", 1))
        args.steps = max(1, args.total_reads // (args.batch_reads * world_))       # batches per rank, all distinct
        args.no_cpu_baseline = True                                                # the scaling run; the CPU arm has its own line
    wl = WORKLOADS[args.workload]
    if args.workload == "tiny":
        args.batch_reads = min(args.batch_reads, 20000)
        args.cpu_sample_pairs = min(args.cpu_sample_pairs, 5000)
    total_bp = sum(s for _, s in wl["chroms"])
    args.workload_desc = (f"{args.workload}: synthetic {len(wl['chroms'])}-chromosome {total_bp / 1e6:.0f} Mb reference "
                          f"(planted repeats, N-padded ends), {wl['read_len']} bp chimeric SMASH reads, min MEM {wl['min_len']}, "
                          f"MAM mode (production flags -rcref -nomap -samin -samout), mappability filter + 50k-style bins")
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        out = run_reference(args, wl, rank, world)
    else:
        out = run_ours(args, wl, rank, world)
    for d in cleanup:
        shutil.rmtree(d, ignore_errors=True)
    if rank == 0 and out is not None:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
        pc = out.get("parity_check")
        if pc is not None and not pc.get("equal"):
            log("PARITY MISMATCH between the reference binary and the GPU path:", pc)
            sys.exit(3)


if __name__ == "__main__":
    main()
