"""The long-frame MFCC (nfft = 1536, model.py:74's configuration) once on the trimmed ragged batch: the launch `ncu` captures."""
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
lr = dspfe.EndpointPlan().detect(pcm, off_d)
mfl = dspfe.MfccPlan(frame_len=480, frame_step=160, nfft=1536, window=np.hamming(480), preemph=0.0, delta_n=3)
mfl.mfcc_delta(pcm, off_d, trim=lr)
torch.cuda.synchronize()
print("ok")
