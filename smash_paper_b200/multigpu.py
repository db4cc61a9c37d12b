"""Read-sharded multi-GPU tail: the host-side exchange around smash_tail_export_keys / phase_a / phase_b.

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  Ranks map contiguous
ranges of read pairs against their own replica of the index; nothing is exchanged on the search/SAM
path.  The tail needs three small exchanges so that the result equals the single-process run bit for
bit:

  1. all_gather of the dupe-set fingerprints {fp1, fp2, ordinal} -> every rank removes pairs whose key
     first appeared on a LOWER rank (smashMEM.py:217-228 is first-wins in name order);
  2. all_gather of the shard edges -> varbin's "same position as the previous kept line" rule
     (varbin.py:56-58) sees the last position of the previous shard;
  3. one all_reduce(sum) of the n_bins int64 counts (+ the five stats counters).

`backend` is any object with export_keys() -> int64 tensor [n,3], phase_a(foreign [m,3]) ->
(n_filtered, first_pos, last_pos), phase_b(has_prev, prev_last_pos) -> (counts tensor, stats dict);
`ContextBackend` wraps a smash_ctx, the CPU tests plug in a numpy model.
"""
from __future__ import annotations

import torch

STAT_KEYS = ["total_reads", "dups_removed", "reads_kept", "n_dupe_pairs", "n_non_dupe_pairs", "n_positions"]


def gather_varlen(t: torch.Tensor, dist, world):
    """all_gather of tensors whose first dimension differs per rank."""
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes + [1])
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[:s] for o, s in zip(out, sizes)]


def lower_rank_keys(parts, rank):
    """Keys of ranks < rank, concatenated (already sorted by ordinal: shards are contiguous)."""
    lower = [p for p in parts[:rank] if p.shape[0]]
    if not lower:
        return parts[rank][:0]
    return torch.cat(lower, dim=0).contiguous()


def previous_last_pos(edges, rank):
    """edges: list of (n_filtered, first_pos, last_pos) per rank -> (has_prev, last position of the nearest
    lower rank that has any filtered position)."""
    for r in range(rank - 1, -1, -1):
        if edges[r][0] > 0:
            return True, int(edges[r][2])
    return False, 0


def _all_to_all_rows(chunks, dist, world):
    """chunks[r] = rows this rank sends to rank r (2-D int64 tensors) -> list of rows received from each rank.
    NCCL: one all_to_all_single with split sizes; gloo (CPU tests): all_gather of the padded buckets."""
    dev = chunks[0].device
    ncol = chunks[0].shape[1]
    send_counts = torch.tensor([c.shape[0] for c in chunks], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(send_counts) for _ in range(world)]
    dist.all_gather(all_counts, send_counts)                       # all_counts[src][dst]
    rank = dist.get_rank()
    recv_counts = [int(all_counts[src][rank].item()) for src in range(world)]
    send = torch.cat(chunks, dim=0).contiguous() if sum(c.shape[0] for c in chunks) else torch.zeros((0, ncol), dtype=torch.int64, device=dev)
    if dist.get_backend() == "nccl":
        recv = torch.zeros((sum(recv_counts), ncol), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=[int(c.shape[0]) for c in chunks])
        out, o = [], 0
        for n in recv_counts:
            out.append(recv[o:o + n]); o += n
        return out
    everything = gather_varlen(send, dist, world)                  # CPU fallback: everyone sees every bucket
    out = []
    for src in range(world):
        start = int(sum(int(all_counts[src][d].item()) for d in range(rank)))
        out.append(everything[src][start:start + recv_counts[src]])
    return out


def partitioned_min_ordinals(keys, dist, rank, world):
    """For every row {fp1, fp2, ordinal} of `keys` the smallest ordinal over ALL ranks that carries the same
    (fp1, fp2).  The key space is hash-partitioned: rank r resolves the keys with fp1 mod world == r, so the
    work and the traffic per rank stay constant as ranks are added (weak scaling), unlike an all_gather."""
    import os, time
    dbg = os.environ.get("SMASH_DEBUG_TIMING") and rank == world - 1
    tl = [time.perf_counter()]

    def lap(label):
        if dbg:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            tl.append(time.perf_counter())
            print(f"[smash-dbg]   partitioned:{label:14s} {1e3 * (tl[-1] - tl[-2]):8.3f} ms", flush=True)

    n = keys.shape[0]
    dev = keys.device
    owner = torch.remainder(keys[:, 0], world) if n else torch.zeros(0, dtype=torch.int64, device=dev)
    order = torch.argsort(owner, stable=True)
    sorted_keys = keys[order]
    counts = torch.bincount(owner, minlength=world).tolist() if n else [0] * world
    chunks, o = [], 0
    for c in counts:
        chunks.append(sorted_keys[o:o + c]); o += c
    lap("bucket")
    got = _all_to_all_rows(chunks, dist, world)
    lap("all_to_all")
    recv_counts = [g.shape[0] for g in got]
    allk = torch.cat(got, dim=0) if sum(recv_counts) else torch.zeros((0, 3), dtype=torch.int64, device=dev)
    # group identical (fp1, fp2): three stable sorts (ordinal, fp2, fp1) -> the first row of a group has its minimum ordinal
    if allk.shape[0]:
        p = torch.argsort(allk[:, 2], stable=True)
        p = p[torch.argsort(allk[p, 1], stable=True)]
        p = p[torch.argsort(allk[p, 0], stable=True)]
        s = allk[p]
        head = torch.ones(s.shape[0], dtype=torch.bool, device=dev)
        head[1:] = (s[1:, 0] != s[:-1, 0]) | (s[1:, 1] != s[:-1, 1])
        gid = torch.cumsum(head.to(torch.int64), 0) - 1
        group_min = s[head, 2]
        mins_sorted = group_min[gid]
        mins = torch.empty_like(mins_sorted)
        mins[p] = mins_sorted
    else:
        mins = torch.zeros(0, dtype=torch.int64, device=dev)
    lap("group")
    back_chunks, o = [], 0
    for c in recv_counts:
        back_chunks.append(mins[o:o + c].reshape(-1, 1)); o += c
    back = _all_to_all_rows(back_chunks, dist, world)
    lap("all_to_all 2")
    mins_for_sorted = torch.cat(back, dim=0).reshape(-1) if n else torch.zeros(0, dtype=torch.int64, device=dev)
    result = torch.empty(n, dtype=torch.int64, device=dev)
    result[order] = mins_for_sorted
    return result


def sharded_tail_finish(backend, dist, rank, world):
    """Runs the three exchanges; returns (global counts tensor, global stats dict)."""
    import os
    import time
    dbg = os.environ.get("SMASH_DEBUG_TIMING") and rank == world - 1
    t = [time.perf_counter()]

    def lap(label):
        if dbg:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            t.append(time.perf_counter())
            print(f"[smash-dbg] sharded_finish:{label:14s} {1e3 * (t[-1] - t[-2]):8.3f} ms", flush=True)

    keys = backend.export_keys()
    lap("export")
    if (world > 2 or os.environ.get("SMASH_FORCE_PARTITIONED")) and hasattr(backend, "phase_a_verdict"):
        # many ranks: hash-partitioned exchange, O(1) keys per rank
        mins = partitioned_min_ordinals(keys, dist, rank, world)
        lap("partitioned")
        n_f, first, last = backend.phase_a_verdict(mins)
    else:
        parts = gather_varlen(keys, dist, world) if world > 1 else [keys]
        lap("gather keys")
        foreign = lower_rank_keys(parts, rank)
        lap("cat")
        n_f, first, last = backend.phase_a(foreign)
    lap("phase_a")
    e = torch.tensor([[n_f, first, last]], dtype=torch.int64, device=keys.device)
    edges = [tuple(int(v) for v in x[0]) for x in gather_varlen(e, dist, world)] if world > 1 else [(n_f, first, last)]
    has_prev, prev = previous_last_pos(edges, rank)
    lap("edges")
    counts, stats = backend.phase_b(has_prev, prev)
    lap("phase_b")
    sv = torch.tensor([stats[k] for k in STAT_KEYS], dtype=torch.int64, device=counts.device)
    if world > 1:
        dist.all_reduce(counts)
        dist.all_reduce(sv)
    out = counts, dict(zip(STAT_KEYS, [int(v) for v in sv]))
    lap("allreduce")
    return out


class ContextBackend:
    """smash_ctx behind the backend protocol (CUDA tensors, device pointers straight into the C ABI)."""

    def __init__(self, ctx, ordinal_base, device):
        from . import api
        self.api, self.ctx, self.base, self.device = api, ctx, int(ordinal_base), device
        self.counts = torch.zeros(ctx.n_bins, dtype=torch.int64, device=device)

    def export_keys(self):
        ptr, n = self.ctx.tail_export_keys(self.base)
        t = torch.zeros((n, 3), dtype=torch.int64, device=self.device)
        if n:
            self.api.memcpy(t.data_ptr(), ptr, 24 * n)
        return t

    def phase_a(self, foreign):
        foreign = foreign.contiguous()
        torch.cuda.synchronize(self.device)
        return self.ctx.tail_phase_a(self.base, foreign.data_ptr() if foreign.shape[0] else None, int(foreign.shape[0]))

    def phase_a_verdict(self, min_ord):
        min_ord = min_ord.contiguous()
        torch.cuda.synchronize(self.device)
        return self.ctx.tail_phase_a_verdict(self.base, min_ord.data_ptr() if min_ord.shape[0] else None, int(min_ord.shape[0]))

    def phase_b(self, has_prev, prev):
        _, stats = self.ctx.tail_phase_b(has_prev, prev, self.counts.data_ptr())
        return self.counts, stats
