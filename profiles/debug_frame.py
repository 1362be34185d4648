import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))
import numpy as np, torch
import dspfe
from dspfe import shard, synth
from oracle import ref_features as O
world, rank, U = 2, 0, 4096
uu, ff = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
all_len = synth.ragged_lengths(U * world, seed=2024)
idx = shard.lpt_partition(all_len, world)[rank]
lengths = all_len[idx]
pcm, off = synth.synth_batch_torch(lengths, seed0=555 + 7919 * rank, device=dev)
off_np = off.numpy()
x = pcm[off_np[uu]:off_np[uu + 1]].cpu().numpy()
l, r = O.basic_endpoint_detection(x, 16000)
pre = O.preemphasis(x, 0.97)[l:r]
rows64 = O.pitch_rows_cep(pre, 16000)           # smoothed, float64
sig = O.downsampling(pre, 16000, 10000); fr = O.to_frames(sig, 10000, 0.0512, 0.01)
raw64 = O.pitch_detect_frame(O.center_clip(fr, False), 10000)
xd = torch.from_numpy(x).to(dev); od = torch.tensor([0, len(x)], dtype=torch.int64, device=dev)
trim = torch.tensor([[l, r]], dtype=torch.int32, device=dev)
for RL in (200, 512):
    o = dspfe.PitchPlan(method=0, preemph=0.97, row_len=RL).detect(xd, od, trim=trim, want_rows=True)
    torch.cuda.synchronize()
    F = int(o["frame_off"][-1]); raw = o["rows"][:F].cpu().numpy()
    sm, sc, lg = dspfe.smooth_rows_f32(raw, mode=0, want_score=True, want_lag=True)
    e_raw = np.abs(raw - raw64[:, :RL]); e_sm = np.abs(sm - rows64[:, :RL])
    print("RL", RL, "F", F, "lag gpu", int(o["lag"][ff]), "tap lag", int(lg[ff]), "ref", 20 + int(np.argmax(O.peak_score(rows64[ff]))))
    print("  max|row64|", np.max(np.abs(rows64[ff])), "raw err max (frame)", e_raw[ff].max(), "at", e_raw[ff].argmax(), "smoothed err max", e_sm[ff].max(), "at", e_sm[ff].argmax())
    print("  worst raw err over all frames", e_raw.max(), np.unravel_index(e_raw.argmax(), e_raw.shape), "rel to row max", (e_raw.max(axis=1) / np.max(np.abs(raw64), axis=1)).max())
    w = rows64[ff]; g = sm[ff]
    print("  ref score@63", O.peak_score(w)[43], "gpu score@63", sc[ff][43], "gpu score@39", sc[ff][19], "ref score@39", O.peak_score(w)[19])
    v = w[63]; near = [(j, w[j] - v, g[j] - g[63]) for j in range(63 - 31, 63 + 32) if j != 63 and abs(w[j] - v) < 3e-4 * np.max(np.abs(w))]
    print("  neighbours of lag 63 close in value (j, ref diff, gpu diff):", near[:10])
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dbg_utt.npz"), x=x, lr=np.array([l, r]), raw_gpu=raw, sm_gpu=sm)
