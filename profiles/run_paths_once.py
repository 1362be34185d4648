"""Every front-end path once on one ragged batch (no warm-up, no timing): the launch sequence `ncu` captures for
profiles/rN_paths_ncu.md.  usage: python profiles/run_paths_once.py [U]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))
import numpy as np
import torch
import dspfe
from dspfe import synth

U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
lengths = synth.ragged_lengths(U, seed=33)
pcm, off = synth.synth_batch_torch(lengths, seed0=31337, device=dev)
off_d = off.to(dev)
ep, mf = dspfe.EndpointPlan(), dspfe.MfccPlan(delta_n=2)
cep, acr = dspfe.PitchPlan(method=0, preemph=0.97), dspfe.PitchPlan(method=1, frame_len=300)
acr512 = dspfe.PitchPlan(method=1)
mfl = dspfe.MfccPlan(frame_len=480, frame_step=160, nfft=1536, window=np.hamming(480), preemph=0.0, delta_n=3)
lr = ep.detect(pcm, off_d)
feat, ffo = mf.mfcc_delta(pcm, off_d, trim=lr)
cep.detect(pcm, off_d, trim=lr, want_feat=True)
acr.detect(pcm, off_d, trim=lr)
acr512.detect(pcm, off_d, trim=lr)
dspfe.cmvn_pad_batch(feat, ffo)
ep.detect_robust(pcm, off_d)
mfl.mfcc_delta(pcm, off_d, trim=lr)
torch.cuda.synchronize()
print("ok", U, float(lengths.sum()) / 16000)
