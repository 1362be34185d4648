"""Synthetic 16 kHz int16 speech-like utterances (SURVEY.md §8d recipe).

x = clip(A * env(t) * sum_{k=1..7} sin(k*phi(t))/k + sigma*N(0,1)), f0 in U(80,300) Hz with
<=20 % vibrato, A in U(1000,12000), sigma in U(20,300), env = Hann over a random 40-70 %
interior span, so >=150 ms of noise-only head and tail exists (the endpoint silence model,
reference endpoint.py:151, assumes it).  `synth_utterance` is the NumPy generator used by
tests and goldens; `synth_batch_torch` builds a packed ragged batch on any torch device
(used by bench.py so that large batches never cross PCIe).
"""
import numpy as np


def utterance_params(seed):
    rng = np.random.default_rng(int(seed))
    return dict(
        f0=rng.uniform(80.0, 300.0), vib_depth=rng.uniform(0.0, 0.2), vib_rate=rng.uniform(3.0, 7.0),
        amp=rng.uniform(1000.0, 12000.0), sigma=rng.uniform(20.0, 300.0),
        span=rng.uniform(0.4, 0.7), pos=rng.uniform(0.0, 1.0), noise_seed=int(rng.integers(1 << 31)))


def synth_utterance(seed, n_samples, sr=16000):
    """One utterance as int16[n_samples]."""
    p = utterance_params(seed)
    n = int(n_samples)
    t = np.arange(n, dtype=np.float64) / sr
    # instantaneous frequency f0*(1+d*sin(2*pi*r*t)) integrated to a phase
    phi = 2 * np.pi * p["f0"] * (t - p["vib_depth"] / (2 * np.pi * p["vib_rate"]) * (np.cos(2 * np.pi * p["vib_rate"] * t) - 1))
    voiced = np.zeros(n)
    for k in range(1, 8):
        voiced += np.sin(k * phi) / k
    span = max(int(p["span"] * n), 1)
    margin = min(int(0.15 * sr), max((n - span) // 2, 0))
    start = margin + int(p["pos"] * max(n - span - 2 * margin, 0))
    env = np.zeros(n)
    env[start:start + span] = np.hanning(span)[: max(min(span, n - start), 0)]
    noise = np.random.default_rng(p["noise_seed"]).standard_normal(n) * p["sigma"]
    x = p["amp"] * env * voiced + noise
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def synth_batch(lengths, seed0=0, sr=16000):
    """Packed ragged batch (NumPy): returns (pcm int16[sum], offsets int64[U+1])."""
    lengths = np.asarray(lengths, dtype=np.int64)
    offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    pcm = np.empty(int(offsets[-1]), dtype=np.int16)
    for u, n in enumerate(lengths):
        pcm[offsets[u]:offsets[u + 1]] = synth_utterance(seed0 + u, int(n), sr)
    return pcm, offsets


def ragged_lengths(n_utt, seed=0, lo=8000, hi=80000):
    """S in U{lo..hi} (BASELINE configs 3-5: 0.5-5 s at 16 kHz)."""
    return np.random.default_rng(int(seed)).integers(lo, hi + 1, size=int(n_utt)).astype(np.int64)


def synth_batch_torch(lengths, seed0=0, sr=16000, device="cuda", chunk=256):
    """Same recipe on a torch device.  Returns (pcm int16[sum] tensor, offsets int64[U+1] CPU tensor).
    Not bit-identical to the NumPy generator (different RNG); parity tests always compare on the
    samples actually produced."""
    import torch
    lengths = np.asarray(lengths, dtype=np.int64)
    U = len(lengths)
    offsets = np.zeros(U + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    pcm = torch.empty(int(offsets[-1]), dtype=torch.int16, device=device)
    gen = torch.Generator(device=device)
    for c0 in range(0, U, chunk):
        c1 = min(c0 + chunk, U)
        ln = lengths[c0:c1]
        nmax = int(ln.max())
        ps = [utterance_params(seed0 + u) for u in range(c0, c1)]
        col = lambda k: torch.tensor([p[k] for p in ps], dtype=torch.float32, device=device)[:, None]
        n = torch.tensor(ln, device=device, dtype=torch.float32)[:, None]
        idx = torch.arange(nmax, device=device, dtype=torch.float32)[None, :]
        t = idx / sr
        two_pi = 2 * np.pi
        phi = two_pi * col("f0") * (t - col("vib_depth") / (two_pi * col("vib_rate")) * (torch.cos(two_pi * col("vib_rate") * t) - 1))
        voiced = torch.zeros_like(phi)
        for k in range(1, 8):
            voiced += torch.sin(k * phi) / k
        span = torch.clamp((col("span") * n).floor(), min=1)
        margin = torch.minimum(torch.full_like(n, float(int(0.15 * sr))), torch.clamp(((n - span) / 2).floor(), min=0))
        start = margin + (col("pos") * torch.clamp(n - span - 2 * margin, min=0)).floor()
        rel = (idx - start) / torch.clamp(span - 1, min=1)
        env = torch.where((rel >= 0) & (rel <= 1), 0.5 - 0.5 * torch.cos(two_pi * rel), torch.zeros_like(rel))
        gen.manual_seed(int(seed0) * 1000003 + c0)
        noise = torch.randn(phi.shape, generator=gen, device=device) * col("sigma")
        x = torch.clamp(torch.round(col("amp") * env * voiced + noise), -32768, 32767).to(torch.int16)
        valid = idx < n
        pcm[int(offsets[c0]):int(offsets[c1])] = x[valid]
    return pcm, torch.from_numpy(offsets)
