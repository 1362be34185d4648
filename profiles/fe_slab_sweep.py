"""Slab-size sweep of the whole-front-end entry points on the bench workload (device path and host-buffer path), with a
bit-for-bit comparison of every output against the single-slab device run.  python profiles/fe_slab_sweep.py [U]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dsp-speech-recognition_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import dspfe
from dspfe import synth

U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
lengths = synth.ragged_lengths(U, seed=2024)
pcm, off = synth.synth_batch_torch(lengths, seed0=555, device=dev)
off_np, off_d = off.numpy(), off.to(dev)
total = pcm.numel()


def snap(o, tot):
    r, c, a = tot
    return {k: (v[:r] if k == "mfcc" else v[:c] if k in ("cep_pitch", "cep_lag") else v[:a] if k in ("acr_pitch", "acr_lag") else v).cpu().clone()
            for k, v in o.items()}


def same(a, b):
    return [k for k in a if not torch.equal(torch.nan_to_num(a[k].double()), torch.nan_to_num(b[k].double()))]


ref = None
for slab in (0, 64 << 20, 32 << 20, 16 << 20, 8 << 20):
    fe = dspfe.FrontendPlan(slab_samples=slab)
    o = fe.alloc(total, U, device=dev)
    for _ in range(3):
        tot = fe.run(pcm, off_d, off_np, o)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fe.run(pcm, off_d, off_np, o)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 100
    s = snap(o, tot)
    if ref is None:
        ref = s
    print(f"device path slab_samples={slab >> 20:4d} Mi: {ms:7.3f} ms/step  differing outputs: {same(s, ref)}", flush=True)
    fe.close()

h_pcm = torch.empty(total, dtype=torch.int16).pin_memory()
h_pcm.copy_(pcm)
h_np = h_pcm.numpy()
for slab in (64 << 20, 32 << 20, 16 << 20, 8 << 20, 4 << 20):
    fe = dspfe.FrontendPlan(host_slab_samples=slab)
    o = fe.alloc(total, U, device=None, pinned=True)
    for _ in range(2):
        tot = fe.run_host(h_np, off_np, o)
    t0 = time.perf_counter()
    for _ in range(5):
        fe.run_host(h_np, off_np, o)
    ms = (time.perf_counter() - t0) * 200
    print(f"host path host_slab_samples={slab >> 20:4d} Mi: {ms:7.3f} ms/step  differing outputs: {same(snap(o, tot), ref)}", flush=True)
    fe.close()

# plain copies for scale
d = torch.empty_like(pcm)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    d.copy_(h_pcm, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D of the PCM alone: {(time.perf_counter() - t0) * 200:.3f} ms ({total * 2 / 1e6:.0f} MB)")
