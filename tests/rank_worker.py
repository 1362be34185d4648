"""One rank of tests/test_frontend_gpu.py::test_two_rank_shard_union_equals_single_gpu: processes its LPT shard of the
global utterance list with the whole front-end on cuda:LOCAL_RANK, gathers the MFCC rows over torch.distributed (gloo)
and writes everything to <dir>/rank<r>.npz."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dsp-speech-recognition_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dspfe  # noqa: E402
from dspfe import shard, synth  # noqa: E402


def main():
    U, out_dir = int(sys.argv[1]), sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    lengths = synth.ragged_lengths(U, seed=77, lo=6000, hi=60000)
    idx = shard.shard_for_rank(lengths, rank, world)
    utts = {int(u): synth.synth_utterance(7000 + int(u), int(lengths[u])) for u in idx}
    pcm, off = shard.pack_shard(utts, idx)
    fe = dspfe.FrontendPlan(slab_samples=300000)
    o = fe.alloc(len(pcm), len(idx), device=dev)
    rows, fc, fa = fe.run(torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev), off, o)
    torch.cuda.synchronize()
    fo = o["mfcc_frame_off"].cpu()
    g_rows, g_off = shard.gather_rows(o["mfcc"][:rows].cpu(), np.diff(fo.numpy()), idx, U)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lr=o["lr"].cpu().numpy(), mfcc=o["mfcc"][:rows].cpu().numpy(),
             mfcc_frame_off=fo.numpy(), cep_pitch=o["cep_pitch"][:fc].cpu().numpy(), cep_lag=o["cep_lag"][:fc].cpu().numpy(),
             cep_frame_off=o["cep_frame_off"].cpu().numpy(), cep_feat=o["cep_feat"].cpu().numpy(),
             acr_pitch=o["acr_pitch"][:fa].cpu().numpy(), acr_lag=o["acr_lag"][:fa].cpu().numpy(),
             acr_frame_off=o["acr_frame_off"].cpu().numpy(), gathered_mfcc=g_rows.numpy(), gathered_off=g_off.numpy())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
