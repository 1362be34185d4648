"""BASELINE configs[4] (C5): the full front-end over 1 M synthetic ragged utterances, sharded length-balanced (LPT from ONE
global list) across the ranks of one box.  Launch with torch.distributed.run, one rank per GPU.  Every rank synthesises its
shard on its own GPU (the 88 GB of PCM never cross PCIe), runs dspfe_frontend over it -- the slab loop keeps the workspaces
bounded: at most slab_samples (256 Mi) samples are in flight, i.e. <= 0.5 GB of PCM and <= 9.5 GB of pitch workspace -- and
reports its device time; rank 0 prints one JSON line with per-rank times, the shard imbalance and the aggregate throughput.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/run_c5.py [n_utt] [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))
import numpy as np
import torch
import torch.distributed as dist

import dspfe
from dspfe import shard, synth


def main():
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    all_len = synth.ragged_lengths(n_utt, seed=5)
    t0 = time.perf_counter()
    parts = shard.lpt_partition(all_len, world)
    t_lpt = time.perf_counter() - t0
    idx = parts[rank]
    lengths = all_len[idx]
    t0 = time.perf_counter()
    pcm, off = synth.synth_batch_torch(lengths, seed0=10_000_019 * (rank + 1), device=dev)
    torch.cuda.synchronize()
    t_synth = time.perf_counter() - t0
    off_np, off_d = off.numpy(), off.to(dev)
    fe = dspfe.FrontendPlan(delta_n=2)
    out = fe.alloc(pcm.numel(), len(idx), device=dev)
    tot = fe.run(pcm, off_d, off_np, out)            # warm-up: sizes the workspaces
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tot = fe.run(pcm, off_d, off_np, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    free, total_mem = torch.cuda.mem_get_info()
    # size-independent properties of the result (the parity tests proper run at sizes the oracle can follow)
    fo = out["mfcc_frame_off"]
    ok = bool(int(fo[-1]) == tot[0] and bool((fo[1:] >= fo[:-1]).all()) and bool(torch.isfinite(out["mfcc"][: tot[0]]).all())
              and int(out["cep_lag"][: tot[1]].min()) >= 20 and int(out["cep_lag"][: tot[1]].max()) <= 99
              and int(out["acr_lag"][: tot[2]].min()) >= 20 and int(out["acr_lag"][: tot[2]].max()) <= 199
              and bool((out["lr"][:, 1] >= out["lr"][:, 0]).all()))
    rec = {"rank": rank, "utterances": int(len(idx)), "audio_s": float(lengths.sum()) / 16000, "ms": ms, "pcm_gb": pcm.numel() * 2 / 1e9,
           "rows": list(tot), "synth_s": t_synth, "gpu_mem_used_gb": (total_mem - free) / 1e9, "properties_ok": ok}
    allr = [None] * world
    dist.all_gather_object(allr, rec)
    if rank == 0:
        tmax = max(r["ms"] for r in allr)
        audio = sum(r["audio_s"] for r in allr)
        print(json.dumps({"workload": "configs[4]: full front-end over %d ragged utterances, LPT shards" % n_utt, "n_gpus": world,
                          "audio_s": audio, "ms_max_over_ranks": tmax, "audio_s_per_s": audio / (tmax * 1e-3),
                          "shard_imbalance": shard.imbalance(all_len, parts), "lpt_partition_s": t_lpt,
                          "slab_samples": 256 << 20, "ranks": allr}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
