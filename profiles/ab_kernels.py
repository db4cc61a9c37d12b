#!/usr/bin/env python
"""A/B measurement harness for the per-batch kernels (measurement tooling, not product code).

    python profiles/ab_kernels.py --workload config2 --variants default,vmin8 --l2 128,64,32 --seed-k 0,15 > gpurun_out/ab.jsonl

The parent generates the workload once (reference text + 3 read batches) into /dev/shm; every (variant, seed_k) runs in
its own child process (a variant is another build of the library, `make -C smash_paper_b200/csrc variant NAME=..`,
selected through SMASH_B200_LIB), builds the index on the GPU and then times the resident step under each L2 fetch
granularity (cudaLimitMaxL2FetchGranularity, set through the system libcudart on the shared primary context).
One JSON line per (variant, seed_k, l2): stage milliseconds per 1 M-read step, CUDA events inside the library.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHM = "/dev/shm/smash_ab"
BATCH_KEYS = ("names", "name_off", "seq", "qual", "seq_off", "flags", "opt", "opt_off")


def parent(args):
    import bench
    from smash_paper_b200 import sequence, synth
    wl = bench.WORKLOADS[args.workload]
    os.makedirs(SHM, exist_ok=True)
    t0 = time.time()
    ref = synth.make_reference(wl["chroms"], seed=1, n_families=wl["families"])
    text, startpos, sizes, descr = sequence.text_from_chromosomes(ref.names, ref.seqs, rcref=True)
    np.save(os.path.join(SHM, "text.npy"), text)
    json.dump(dict(startpos=[int(x) for x in startpos], sizes=[int(x) for x in sizes], descr=list(descr), names=ref.names,
                   offsets=[int(x) for x in ref.offsets()], total=int(ref.total)), open(os.path.join(SHM, "meta.json"), "w"))
    genome = ref.concat()
    del text
    for i in range(args.batches):
        b = synth.make_reads_fast(genome, args.batch_reads // 2, read_len=wl["read_len"], seed=1000, first_pair=i * (args.batch_reads // 2))
        np.savez(os.path.join(SHM, f"batch{i}.npz"), **{k: getattr(b, k) for k in BATCH_KEYS})
    starts = bench.make_bins(wl, ref, SHM)
    np.save(os.path.join(SHM, "bins.npy"), starts)
    del genome, ref
    print(f"[ab] workload ready in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    for variant in args.variants.split(","):
        for k in args.seed_k.split(","):
            env = dict(os.environ)
            # "name@VAR=VALUE[@VAR2=..]": the default build with environment switches of the library (e.g. flat@SMASH_FLAT_SEED=1)
            variant, *envs = variant.split("@")
            for kv in envs:
                env[kv.split("=", 1)[0]] = kv.split("=", 1)[1]
            vlib = os.path.join(ROOT, "build", "variants", f"libsmash_b200_{variant}.so")
            if variant != "default" and (os.path.exists(vlib) or not envs):
                env["SMASH_B200_LIB"] = vlib
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--workload", args.workload, "--variant", variant,
                                "--seed-k", k, "--l2", args.l2, "--steps", str(args.steps), "--batches", str(args.batches),
                                "--want", args.want], env=env)
            if r.returncode:
                print(json.dumps({"variant": variant, "seed_k": int(k), "error": r.returncode}), flush=True)
    if not args.keep:
        import shutil
        shutil.rmtree(SHM, ignore_errors=True)


def child(args):
    import bench
    from smash_paper_b200 import api, synth
    wl = bench.WORKLOADS[args.workload]
    meta = json.load(open(os.path.join(SHM, "meta.json")))
    text = np.load(os.path.join(SHM, "text.npy"))
    cudart = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    cudart.cudaSetDevice(0)
    t0 = time.time()
    os.environ["SMASH_L2_FETCH"] = os.environ.get("AB_L2_FETCH", "0")     # 0: the library leaves the limit alone; this script sets it
    ctx = api.Context.from_text(text, meta["startpos"], meta["sizes"], meta["descr"], keep_isa=True, chunk_cap=wl["chunk_cap"],
                                min_len=wl["min_len"], nomap=True, tag_mappability=True, seed_k=int(args.seed_k))
    n_text = len(text)
    del text
    ctx.build_mappability_device()
    ctx.drop_isa()
    t_index = time.time() - t0
    ctx.tail_configure(np.load(os.path.join(SHM, "bins.npy")), meta["names"], meta["offsets"])
    batches = []
    for i in range(args.batches):
        z = np.load(os.path.join(SHM, f"batch{i}.npz"))
        batches.append(synth.ReadBatch(**{k: z[k] for k in BATCH_KEYS}))
    want = {"sam_tail": api.WANT_SAM | api.WANT_TAIL, "tail": api.WANT_TAIL, "sam": api.WANT_SAM}[args.want]
    ctx.tail_reserve((batches[0].n // 2) * (args.steps + 3) * 4, 8 * batches[0].n * (args.steps + 3))
    for l2 in args.l2.split(","):
        l2 = int(l2)
        if l2:
            rc = cudart.cudaDeviceSetLimit(ctypes.c_int(5), ctypes.c_size_t(l2))         # cudaLimitMaxL2FetchGranularity
            got = ctypes.c_size_t(0)
            cudart.cudaDeviceGetLimit(ctypes.byref(got), ctypes.c_int(5))
        else:
            rc, got = 0, ctypes.c_size_t(0)
        for i in range(2):
            ctx.upload(batches[i % len(batches)]); ctx.map_resident(want)
        ctx.tail_reset(); ctx.stage_ms(reset=True)
        ms = 0.0
        for i in range(args.steps):
            ctx.upload(batches[i % len(batches)])
            ms += ctx.map_resident(want).gpu_ms
        st = ctx.stage_ms(reset=True)
        ctx.tail_reset()
        print(json.dumps({"variant": args.variant, "seed_k": int(args.seed_k), "l2_req": l2, "l2_rc": rc, "l2_got": int(got.value),
                          "ms_per_step": ms / args.steps, "stage_ms": {k: round(v / args.steps, 4) for k, v in st.items()},
                          "index_s": round(t_index, 1), "index_gb": round(ctx.index_bytes / 1e9, 2), "n_text": n_text, "want": args.want}), flush=True)
    ctx.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--variants", default="default")
    ap.add_argument("--variant", default="default")
    ap.add_argument("--seed-k", default="0")
    ap.add_argument("--l2", default="0")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--batch-reads", type=int, default=1_000_000)
    ap.add_argument("--want", default="sam_tail")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    child(a) if a.child else parent(a)
