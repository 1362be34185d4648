"""The bench workload's front-end step, twice (one warm pass, one for the profiler), no timing: the launch sequence `ncu`
captures for profiles/r2_*.  usage: python profiles/run_frontend_once.py [U]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))
import torch
import dspfe
from dspfe import synth

U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
lengths = synth.ragged_lengths(U, seed=2024)
pcm, off = synth.synth_batch_torch(lengths, seed0=555, device=dev)
fe = dspfe.FrontendPlan(delta_n=2)
o = fe.alloc(pcm.numel(), U, device=dev)
for _ in range(2):
    tot = fe.run(pcm, off.to(dev), off.numpy(), o)
torch.cuda.synchronize()
print("ok", U, float(lengths.sum()) / 16000, tot)
