"""Debug aid: rank `r` of `world`'s bench shard on one GPU; device path vs host path, and the parity sample vs the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))
import numpy as np, torch
import dspfe
from dspfe import shard, synth
from oracle import ref_features as O
world, rank, U = int(sys.argv[1]), int(sys.argv[2]), 4096
dev = torch.device("cuda:0")
all_len = synth.ragged_lengths(U * world, seed=2024)
idx = shard.lpt_partition(all_len, world)[rank]
lengths = all_len[idx]
pcm, off = synth.synth_batch_torch(lengths, seed0=555 + 7919 * rank, device=dev)
off_np, off_d = off.numpy(), off.to(dev)
fe = dspfe.FrontendPlan(delta_n=2)
out = fe.alloc(pcm.numel(), len(idx), device=dev)
tot = fe.run(pcm, off_d, off_np, out); torch.cuda.synchronize()
h_pcm = pcm.cpu().numpy()
h_out = fe.alloc(pcm.numel(), len(idx), device=None, pinned=True)
for rep in range(3):
    htot = fe.run_host(h_pcm, off_np, h_out)
    print("rep", rep, "totals", tot, htot)
    for k in ("lr", "mfcc", "cep_lag", "acr_lag", "cep_pitch", "acr_pitch", "mfcc_frame_off", "cep_frame_off", "acr_frame_off", "cep_feat"):
        n = {"mfcc": tot[0], "cep_lag": tot[1], "cep_pitch": tot[1], "acr_lag": tot[2], "acr_pitch": tot[2]}.get(k, None)
        a, b = out[k].cpu(), h_out[k]
        if n is not None: a, b = a[:n], b[:n]
        eq = torch.equal(torch.nan_to_num(a.double()), torch.nan_to_num(b.double()))
        if not eq:
            d = (torch.nan_to_num(a.double()) != torch.nan_to_num(b.double())).reshape(len(a), -1).any(1).nonzero().flatten()
            print("  DIFF", k, len(d), d[:10].tolist())
# oracle check of the first 128 utterances
fo_c = out["cep_frame_off"].cpu().numpy(); fo_a = out["acr_frame_off"].cpu().numpy(); lr = out["lr"].cpu().numpy()
cl = out["cep_lag"].cpu().numpy(); al = out["acr_lag"].cpu().numpy()
for u in range(128):
    x = h_pcm[off_np[u]:off_np[u + 1]]
    l, r = int(lr[u, 0]), int(lr[u, 1])
    assert (l, r) == O.basic_endpoint_detection(x, 16000)
    pre = O.preemphasis(x, 0.97)[l:r]
    rows = O.pitch_rows_cep(pre, 16000)
    lag = np.array([20 + int(np.argmax(O.peak_score(c))) for c in rows])
    g = cl[fo_c[u]:fo_c[u + 1]]
    for i in np.nonzero(lag != g)[0]:
        lo, hi = O.peak_score_bounds(rows[i], 1e-5 * np.max(np.abs(rows[i])))
        sc = np.array(O.peak_score(rows[i]))
        print("cep utt", u, "frame", i, "of", len(lag), "gpu", g[i], "ref", lag[i], "near", O.lag_is_near_tie_cep(rows[i], g[i]), "score ref", sc[lag[i]-20], "score at gpu lag", sc[g[i]-20], "lo/hi at gpu", lo[g[i]-20], hi[g[i]-20], "lo.max", lo.max())
    srows = np.asarray(O.pitch_scores_sr(x[l:r], 16000, winlen=0.03, step=0.01)[0])
    slag = 20 + np.argmax(srows, axis=1)
    g = al[fo_a[u]:fo_a[u + 1]]
    for i in np.nonzero(slag != g)[0]:
        rw = srows[i]
        print("acr utt", u, "frame", i, "of", len(slag), "gpu", g[i], "ref", slag[i], "near", O.lag_is_near_tie_sr(rw, g[i]), "vals", rw[g[i]-20], rw[slag[i]-20], "rel", (rw[slag[i]-20]-rw[g[i]-20])/np.max(np.abs(rw)))
print("done")
