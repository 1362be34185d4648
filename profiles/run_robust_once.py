"""robust_endpoint_detection (K3r) once on the ragged batch: the launch `ncu` captures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))
import torch
import dspfe
from dspfe import synth

U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
lengths = synth.ragged_lengths(U, seed=33)
pcm, off = synth.synth_batch_torch(lengths, seed0=31337, device=dev)
off_d = off.to(dev)
ep = dspfe.EndpointPlan()
for _ in range(2):
    lr = ep.detect_robust(pcm, off_d)
torch.cuda.synchronize()
print("ok")
