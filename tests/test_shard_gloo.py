"""Multi-GPU host logic on CPU: the LPT partition and the optional final gather over gloo, world_size 2."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lengths(n=300, seed=1):
    return np.random.default_rng(seed).integers(8000, 80001, size=n)


def test_lpt_partition_is_a_balanced_partition():
    from dspfe import shard
    ln = _lengths(1000)
    for world in (1, 2, 4, 8):
        parts = shard.lpt_partition(ln, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(len(ln)))                 # a partition: every utterance exactly once
        assert all(np.all(np.diff(p) > 0) for p in parts if len(p) > 1)          # original order kept inside a shard
        loads = [ln[p].sum() for p in parts]
        assert max(loads) - min(loads) <= ln.max()                              # LPT bound: within one utterance
        assert shard.imbalance(ln, parts) < 1.01
        assert all(np.array_equal(a, b) for a, b in zip(parts, shard.lpt_partition(ln, world)))   # deterministic


def _worker(rank, world, port, tmp):
    for p in (ROOT, os.path.join(ROOT, "dsp-speech-recognition_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dspfe import shard
    ln = _lengths(57, seed=9)
    idx = shard.shard_for_rank(ln, rank, world)
    # stand-in for the per-utterance feature rows a rank computed: frame count = len // 160, row value = utterance id
    counts = ln[idx] // 160
    rows = torch.cat([torch.full((int(c), 3), float(u)) for u, c in zip(idx, counts)])
    out, off = shard.gather_rows(rows, counts, idx, len(ln))
    want_counts = ln // 160
    assert np.array_equal(np.diff(off.numpy()), want_counts)
    for u in range(len(ln)):
        blk = out[int(off[u]): int(off[u + 1])]
        assert blk.shape[0] == want_counts[u] and bool((blk == float(u)).all())
    # weak-scaling accounting used by bench.py: sum over ranks of shard audio == total audio
    t = torch.tensor([float(ln[idx].sum())], dtype=torch.float64)
    dist.all_reduce(t)
    assert float(t) == float(ln.sum())
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_gloo_world2_shards_and_gather(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))
