"""Per-path device timings (CUDA events) of the front-end on one GPU: endpoint, MFCC, cepstrum pitch (+feature),
autocorrelation pitch on a ragged batch (BASELINE configs 3/4).  usage: python profiles/time_paths.py [U] [iters]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))
import numpy as np
import torch
import dspfe
from dspfe import synth

U = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
lengths = synth.ragged_lengths(U, seed=33)
pcm, off = synth.synth_batch_torch(lengths, seed0=31337, device=dev)
off_d = off.to(dev)
audio_s = float(lengths.sum()) / 16000
ep, mf = dspfe.EndpointPlan(), dspfe.MfccPlan(delta_n=2)
cep, acr = dspfe.PitchPlan(method=0, preemph=0.97), dspfe.PitchPlan(method=1, frame_len=300)
lr = ep.detect(pcm, off_d)
bufs = {}


def timed(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(json.dumps({"path": name, "ms": ms, "audio_s_per_s": audio_s / (ms * 1e-3), "utterances": U, "audio_s": audio_s}))


out = torch.empty((mf.rows_bound(pcm.numel(), U), 39), dtype=torch.float32, device=dev)
fo = torch.empty(U + 1, dtype=torch.int64, device=dev)
timed("endpoint", lambda: ep.detect(pcm, off_d))
timed("mfcc_delta (trimmed)", lambda: mf.mfcc_delta(pcm, off_d, trim=lr, out=out, frame_off=fo))
o1 = cep.detect(pcm, off_d, trim=lr, want_feat=True)
timed("pitch_cepstrum + pitch_feature", lambda: cep.detect(pcm, off_d, trim=lr, want_feat=True, out=o1))
o2 = acr.detect(pcm, off_d, trim=lr)
timed("pitch_autocorrelation (trimmed, 300-sample frames)", lambda: acr.detect(pcm, off_d, trim=lr, out=o2))

# ---- rows SURVEY section 8 marks "next": the trainer's batching epilogue (f-1), the gated endpoint rule (f-3), WAV ingest (f-4)
feat, ffo = mf.mfcc_delta(pcm, off_d, trim=lr, out=out, frame_off=fo)
inp = torch.empty((200, U, 39), dtype=torch.float32, device=dev)
timed("cmvn_pad_batch (f-1: CMVN + pad to [200,U,39])", lambda: dspfe.cmvn_pad_batch(feat, ffo, out=inp))
timed("robust_endpoint_detection (f-3: autocorrelation-gated rule)", lambda: ep.detect_robust(pcm, off_d))
# f-2: model.py:74's configuration (30 ms Hamming frames, nfft 1536) on the trimmed batch
mfl = dspfe.MfccPlan(frame_len=480, frame_step=160, nfft=1536, window=np.hamming(480), preemph=0.0, delta_n=3)
outl = torch.empty((mfl.rows_bound(pcm.numel(), U), 39), dtype=torch.float32, device=dev)
timed("mfcc_delta nfft=1536 (f-2: 480-sample Hamming frames, trimmed)", lambda: mfl.mfcc_delta(pcm, off_d, trim=lr, out=outl, frame_off=fo))
# the same call at the data set's own rate (report p.9: 44.1 kHz -> 1323-sample frames, hop 441); the batch is read as 44.1 kHz audio
mfh = dspfe.MfccPlan(frame_len=1323, frame_step=441, nfft=1536, window=np.hamming(1323), preemph=0.0, delta_n=3, samplerate=44100)
outh = torch.empty((mfh.rows_bound(pcm.numel(), U), 39), dtype=torch.float32, device=dev)
timed("mfcc_delta nfft=1536 (f-2: 1323-sample Hamming frames, hop 441, untrimmed)", lambda: mfh.mfcc_delta(pcm, off_d, out=outh, frame_off=fo))
if "--ingest" in sys.argv:
    import tempfile, time
    from scipy.io import wavfile
    d = tempfile.mkdtemp()
    paths = []
    rng = np.random.default_rng(0)
    for i in range(256):
        x = rng.integers(-3000, 3000, size=(int(lengths[i]), 2), dtype=np.int16)
        paths.append(os.path.join(d, f"u{i}.wav")); wavfile.write(paths[-1], 16000, x)
    dspfe.ingest_wavs(paths[:8])
    t0 = time.perf_counter(); p2, o2_, _ = dspfe.ingest_wavs(paths); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"path": "ingest_wavs (f-4: 256 stereo files from the page cache)", "ms": dt * 1e3,
                      "audio_s_per_s": float(o2_[-1]) / 16000 / dt, "file_MB_per_s": sum(os.path.getsize(p) for p in paths) / 1e6 / dt}))
