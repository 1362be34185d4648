"""Length-balanced sharding of a ragged utterance list across the GPUs of one box (BASELINE config 5).

Utterances are independent, so the hot path needs no collective: every rank computes the same deterministic
partition from the (global) length list, packs and processes its own shard, and keeps its features (a
data-parallel consumer reads them where they are).  `gather_rows` is the optional final feature gather over
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests); it is off the hot path.
"""
import heapq

import numpy as np


def lpt_partition(lengths, world):
    """Longest-processing-time-first: sort by length descending, give each utterance to the least-loaded rank.
    Returns a list of `world` int64 index arrays (each ascending, so a shard keeps the original utterance order).
    Deterministic: ties break on the utterance index, then on the rank."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.lexsort((np.arange(len(lengths)), -lengths))
    heap = [(0, r) for r in range(int(world))]
    buckets = [[] for _ in range(int(world))]
    for u in order:
        load, r = heapq.heappop(heap)
        buckets[r].append(int(u))
        heapq.heappush(heap, (load + int(lengths[u]), r))
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def shard_for_rank(lengths, rank, world):
    """Indices of the utterances rank `rank` owns."""
    return lpt_partition(lengths, world)[int(rank)]


def imbalance(lengths, parts):
    """max shard load / mean shard load (1.0 = perfect)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    loads = np.array([lengths[p].sum() for p in parts], dtype=np.float64)
    return float(loads.max() / max(loads.mean(), 1.0))


def pack_shard(utterances, idx):
    """Packs the utterances `idx` of a list of 1-D int16 arrays: (pcm int16[sum], offsets int64[len(idx)+1])."""
    off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum([len(utterances[i]) for i in idx], out=off[1:])
    pcm = np.empty(int(off[-1]), dtype=np.int16)
    for k, i in enumerate(idx):
        pcm[off[k]:off[k + 1]] = utterances[i]
    return pcm, off


def gather_rows(rows, counts, idx, n_total, group=None):
    """Optional final gather.  `rows` [sum(counts), W] are this rank's per-utterance row blocks (in shard order),
    `counts` [len(idx)] their row counts, `idx` the global utterance indices of the shard.  Returns, on every rank,
    (all_rows [sum over all utterances, W] in global utterance order, offsets int64 [n_total+1])."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = rows.device
    counts_t = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=dev)
    idx_t = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=dev)
    # shard sizes differ: exchange them, then pad every contribution to the largest
    sizes = torch.tensor([idx_t.numel(), rows.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    max_u = int(max(int(s[0]) for s in all_sizes)); max_r = int(max(int(s[1]) for s in all_sizes))
    pad_i = torch.full((max_u,), -1, dtype=torch.int64, device=dev); pad_i[: idx_t.numel()] = idx_t
    pad_c = torch.zeros((max_u,), dtype=torch.int64, device=dev); pad_c[: idx_t.numel()] = counts_t
    pad_r = torch.zeros((max_r, rows.shape[1]), dtype=rows.dtype, device=dev); pad_r[: rows.shape[0]] = rows
    g_i = [torch.empty_like(pad_i) for _ in range(world)]
    g_c = [torch.empty_like(pad_c) for _ in range(world)]
    g_r = [torch.empty_like(pad_r) for _ in range(world)]
    dist.all_gather(g_i, pad_i, group=group); dist.all_gather(g_c, pad_c, group=group); dist.all_gather(g_r, pad_r, group=group)
    cnt = torch.zeros(n_total, dtype=torch.int64, device=dev)
    n_u = [int(s[0]) for s in all_sizes]; n_r = [int(s[1]) for s in all_sizes]
    for r in range(world):
        cnt[g_i[r][:n_u[r]]] = g_c[r][:n_u[r]]
    off = torch.zeros(n_total + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(cnt, 0)
    out = torch.empty((int(off[-1]), rows.shape[1]), dtype=rows.dtype, device=dev)
    for r in range(world):
        # destination of every source row of rank r: its utterance's first global row + its position inside the block
        c, u = g_c[r][:n_u[r]], g_i[r][:n_u[r]]
        src_start = torch.cumsum(c, 0) - c
        dst = torch.repeat_interleave(off[u] - src_start, c) + torch.arange(n_r[r], device=dev)
        out.index_copy_(0, dst, g_r[r][:n_r[r]])
    return out, off
