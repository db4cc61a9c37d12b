"""CPU, world_size 2 over gloo: the host-side exchange of the read-sharded tail
(smash_paper_b200/multigpu.py) with a numpy model of the per-rank tail phases, against the
single-process result.  The CUDA implementation of the same phases is checked on one GPU in
tests/test_gpu_parity.py::test_two_shards_equal_one_run."""
import bisect
import hashlib
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smash_paper_b200 import multigpu

N_BINS = 64
STARTS = [i * 1000 for i in range(N_BINS)]


def make_pairs(seed=3, n=400):
    """Per-pair kept hits [(tid,pos)...] with duplicate keys within and across shards, empty pairs,
    and equal consecutive positions straddling the shard boundary."""
    rng = np.random.default_rng(seed)
    pairs = []
    for i in range(n):
        k = int(rng.integers(0, 4))
        pairs.append([(int(rng.integers(0, 3)), int(rng.integers(0, 60000))) for _ in range(k)])
    for i in range(30, n, 37):                       # duplicates of an earlier pair (any distance)
        j = int(rng.integers(0, i))
        pairs[i] = list(pairs[j])
    half = n // 2
    pairs[half - 1] = [(0, 1234), (1, 777)]          # last kept line of shard 0 ...
    pairs[half] = [(2, 777), (0, 999)]               # ... first of shard 1 has the same position string
    pairs[half + 1] = []
    return pairs


def fingerprint(hits):
    d = hashlib.blake2b(repr(hits).encode(), digest_size=16).digest()
    return int.from_bytes(d[:8], "little", signed=True), int.from_bytes(d[8:], "little", signed=True)


def single_process(pairs):
    seen, F = set(), []
    nd = nn = 0
    for h in pairs:
        if not h:
            continue
        key = tuple(h)
        if key in seen:
            nd += 1
            continue
        seen.add(key); nn += 1
        F.extend(p for _, p in h)
    counts = [0] * N_BINS
    dups = 0
    prev = None
    for p in F:
        if p == prev:
            dups += 1
            continue
        counts[bisect.bisect(STARTS, p) - 1] += 1
        prev = p
    return counts, dict(total_reads=len(F), dups_removed=dups, reads_kept=len(F) - dups, n_dupe_pairs=nd, n_non_dupe_pairs=nn,
                        n_positions=len(F))


class NumpyTail:
    def __init__(self, pairs, base):
        self.pairs, self.base = pairs, base

    def export_keys(self):
        rows = [fingerprint(h) + (self.base + i,) for i, h in enumerate(self.pairs) if h]
        return torch.tensor(rows, dtype=torch.int64).reshape(-1, 3)

    def phase_a(self, foreign):
        first = {}
        for a, b, o in foreign.tolist():
            first[(a, b)] = min(o, first.get((a, b), o))
        self.F, self.nd, self.nn = [], 0, 0
        for i, h in enumerate(self.pairs):
            if not h:
                continue
            k = fingerprint(h)
            if k in first and first[k] < self.base + i:
                self.nd += 1
                continue
            first.setdefault(k, self.base + i)
            self.nn += 1
            self.F.extend(p for _, p in h)
        return len(self.F), (self.F[0] if self.F else 0), (self.F[-1] if self.F else 0)

    def phase_a_verdict(self, min_ord):
        mo = min_ord.tolist()
        self.F, self.nd, self.nn = [], 0, 0
        j = 0
        for i, h in enumerate(self.pairs):
            if not h:
                continue
            if mo[j] == self.base + i:
                self.nn += 1
                self.F.extend(p for _, p in h)
            else:
                self.nd += 1
            j += 1
        return len(self.F), (self.F[0] if self.F else 0), (self.F[-1] if self.F else 0)

    def phase_b(self, has_prev, prev):
        counts = torch.zeros(N_BINS, dtype=torch.int64)
        dups = 0
        last = prev if has_prev else None
        for p in self.F:
            if p == last:
                dups += 1
                continue
            counts[bisect.bisect(STARTS, p) - 1] += 1
            last = p
        return counts, dict(total_reads=len(self.F), dups_removed=dups, reads_kept=len(self.F) - dups, n_dupe_pairs=self.nd,
                            n_non_dupe_pairs=self.nn, n_positions=len(self.F))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = make_pairs()
    n = len(pairs)
    lo, hi = n * rank // world, n * (rank + 1) // world
    counts, stats = multigpu.sharded_tail_finish(NumpyTail(pairs[lo:hi], lo), dist, rank, world)
    if rank == 0:
        q.put((counts.tolist(), stats))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_equal_single_process(world):
    """world 2: all_gather of the keys of lower ranks; world 3: hash-partitioned exchange + verdict."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, stats = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_counts, exp_stats = single_process(make_pairs())
    assert counts == exp_counts
    assert stats == exp_stats
    assert exp_stats["n_dupe_pairs"] > 5 and exp_stats["dups_removed"] >= 1     # the cases are exercised


def test_helpers():
    assert multigpu.previous_last_pos([(0, 0, 0), (3, 5, 9), (0, 0, 0), (2, 1, 1)], 3) == (True, 9)
    assert multigpu.previous_last_pos([(0, 0, 0), (3, 5, 9)], 1) == (False, 0)
    parts = [torch.tensor([[1, 2, 0]]), torch.zeros((0, 3), dtype=torch.int64), torch.tensor([[3, 4, 7]])]
    assert multigpu.lower_rank_keys(parts, 2).tolist() == [[1, 2, 0]]
    assert multigpu.lower_rank_keys(parts, 0).shape[0] == 0
