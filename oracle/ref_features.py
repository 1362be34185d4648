"""CPU oracle: float64 NumPy restatement of the reference `features` package.

TEST INFRASTRUCTURE ONLY.  Nothing under `dsp-speech-recognition_b200/` may import
this module; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do, and only as the checker / CPU baseline.

Every function below restates one function of AuCson/DSP-Speech-Recognition
(`/root/reference/features/*.py`) and cites the file:line it follows.  The
restatement is vectorised over frames (the reference loops in Python) but keeps
the reference's arithmetic, its container types where callers depend on them,
and its quirks.

Parity pin: the reference ships no golden vectors or tests, so
this oracle is pinned against the *live* reference functions, imported from
/root/reference in the CPU container by `oracle/validate_against_reference.py`
and `tests/golden/make_golden.py`; the outputs of the reference itself are
committed as `tests/golden/*.npz` and `tests/test_oracle_golden.py` checks the
oracle against them on every run.
"""
from __future__ import annotations

import decimal
import math

import numpy as np

EPS64 = float(np.finfo(np.float64).eps)  # base.py:26,30 floor

# Endpoint framing is global config in the reference (config.py:31-32).
CFG_FRAME = 0.03
CFG_STEP = 0.01


# --------------------------------------------------------------------------
# sigproc.py
# --------------------------------------------------------------------------
def round_half_up(number):
    """sigproc.py:55-56."""
    return int(decimal.Decimal(number).quantize(decimal.Decimal("1"), rounding=decimal.ROUND_HALF_UP))


def num_frames(slen, frame_len, frame_step):
    """Frame count rule of framesig, sigproc.py:76-82."""
    if slen <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * slen - frame_len) / frame_step))


def framesig(sig, frame_len, frame_step, winfunc=lambda n: np.ones((n,))):
    """sigproc.py:66-98: zero-pad to (F-1)*step+len, overlapping frames, times window."""
    sig = np.asarray(sig)
    slen = len(sig)
    frame_len = int(round_half_up(frame_len))
    frame_step = int(round_half_up(frame_step))
    nf = num_frames(slen, frame_len, frame_step)
    padlen = (nf - 1) * frame_step + frame_len
    pad = np.concatenate((sig, np.zeros((padlen - slen,))))
    idx = np.arange(frame_len)[None, :] + frame_step * np.arange(nf)[:, None]
    return pad[idx] * winfunc(frame_len)


def to_frames(sig, rate, t=0.020, step=0.010):
    """sigproc.py:11-19: truncating int() of the products before framing."""
    return framesig(sig, int(rate * t), int(step * rate))


def deframesig(frames, siglen, frame_len, frame_step, winfunc=lambda n: np.ones((n,))):
    """sigproc.py:101-133: overlap-add inverse of framesig."""
    frame_len = round_half_up(frame_len)
    frame_step = round_half_up(frame_step)
    frames = np.asarray(frames)
    nf = frames.shape[0]
    assert frames.shape[1] == frame_len, '"frames" matrix is wrong size, 2nd dim is not equal to frame_len'
    padlen = (nf - 1) * frame_step + frame_len
    if siglen <= 0:
        siglen = padlen
    rec = np.zeros((padlen,))
    corr = np.zeros((padlen,))
    win = winfunc(frame_len)
    for i in range(nf):
        sl = slice(i * frame_step, i * frame_step + frame_len)
        corr[sl] = corr[sl] + win + 1e-15
        rec[sl] = rec[sl] + frames[i]
    return (rec / corr)[0:siglen]


def magspec(frames, NFFT):
    """sigproc.py:136-148 (rfft truncates frames longer than NFFT)."""
    return np.absolute(np.fft.rfft(frames, NFFT))


def powspec(frames, NFFT):
    """sigproc.py:151-158."""
    return 1.0 / NFFT * np.square(magspec(frames, NFFT))


def logpowspec(frames, NFFT, norm=1):
    """sigproc.py:161-175."""
    ps = powspec(frames, NFFT)
    ps[ps <= 1e-30] = 1e-30
    lps = 10 * np.log10(ps)
    return lps - np.max(lps) if norm else lps


def preemphasis(signal, coeff=0.95):
    """sigproc.py:178-185 == preprocess.py:11-19.

    For a 2-D (1,S) input `signal[0]` is the whole row and `signal[1:]` is empty,
    so the result is the unfiltered flattened row (Appendix A-1)."""
    signal = np.asarray(signal)
    return np.append(signal[0], signal[1:] - coeff * signal[:-1])


def fir_taps(N, rate, low_freq=0, high_freq=500, wintype="square"):
    """Taps of the one-sided ideal band-pass of `window`, sigproc.py:33-44."""
    Hd = np.zeros(N)
    Hd[int(N * low_freq / rate):int(N * high_freq / rate)] = 1
    hd = np.fft.ifft(Hd, N)
    w = np.hamming(N) if wintype == "hamming" else np.ones(N)
    return 2 * np.pi * w * hd


def window(sig, rate, low_freq=0, high_freq=500, wintype="square"):
    """sigproc.py:22-46: causal complex FIR, output truncated to len(sig)."""
    sig = np.asarray(sig)
    N = sig.shape[-1]
    h = fir_taps(N, rate, low_freq, high_freq, wintype)
    if sig.ndim == 1:
        return np.convolve(sig, h)[:N]
    # batched: y[f, n] = sum_{m<=n} sig[f, m] * h[n-m], the first N samples of the linear convolution, evaluated through a
    # zero-padded 2N-point transform (equal to np.convolve to ~1e-16 relative; validated against the live reference)
    return np.fft.ifft(np.fft.fft(sig, 2 * N, axis=-1) * np.fft.fft(h, 2 * N), axis=-1)[..., :N]


def acr(frame, n):
    """sigproc.py:48-53: unbiased autocorrelation at lag n."""
    frame = np.asarray(frame)
    if n == 0:
        return np.sum(frame * frame) / len(frame)
    return np.sum(frame[:-n] * frame[n:]) / (len(frame) - n)


# --------------------------------------------------------------------------
# base.py
# --------------------------------------------------------------------------
def hz2mel(hz):
    """base.py:34-35."""
    return 2595 * np.log10(1 + hz / 700.0)


def mel2hz(mel):
    """base.py:37-38."""
    return 700 * (10 ** (mel / 2595.0) - 1)


def get_filterbanks(nfilt=20, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    """base.py:40-58."""
    highfreq = highfreq or samplerate / 2
    assert highfreq <= samplerate / 2, "highfreq is greater than samplerate/2"
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    bins = np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        lo, mid, hi = bins[j], bins[j + 1], bins[j + 2]
        i = np.arange(int(lo), int(mid))
        fb[j, i] = (i - lo) / (mid - lo)
        i = np.arange(int(mid), int(hi))
        fb[j, i] = (hi - i) / (hi - mid)
    return fb


def fbank(signal, samplerate=16000, winlen=0.025, winstep=0.01, nfilt=26, nfft=512, lowfreq=0,
          highfreq=None, preemph=0.97, winfunc=lambda n: np.ones((n,))):
    """base.py:18-32."""
    highfreq = highfreq or samplerate / 2
    signal = preemphasis(signal, preemph)
    frames = framesig(signal, winlen * samplerate, winstep * samplerate, winfunc)
    pspec = powspec(frames, nfft)
    energy = np.sum(pspec, 1)
    energy = np.where(energy == 0, EPS64, energy)
    fb = get_filterbanks(nfilt, nfft, samplerate, lowfreq, highfreq)
    feat = np.dot(pspec, fb.T)
    feat = np.where(feat == 0, EPS64, feat)
    return feat, energy


def dct2_ortho(x, numcep):
    """scipy.fftpack.dct(type=2, axis=1, norm='ortho')[:, :numcep] as a matrix product (base.py:13)."""
    n = x.shape[1]
    k = np.arange(numcep)[:, None]
    m = np.arange(n)[None, :]
    C = np.cos(np.pi * k * (2 * m + 1) / (2.0 * n)) * np.sqrt(2.0 / n)
    C[0] *= np.sqrt(0.5)
    return x @ C.T


def lifter(cepstra, L=22):
    """base.py:60-68."""
    if L > 0:
        ncoeff = np.shape(cepstra)[1]
        lift = 1 + (L / 2.0) * np.sin(np.pi * np.arange(ncoeff) / L)
        return lift * cepstra
    return cepstra


def mfcc(signal, samplerate=16000, winlen=0.025, winstep=0.01, numcep=13, nfilt=26, nfft=512,
         lowfreq=0, highfreq=None, preemph=0.97, ceplifter=22, appendEnergy=True,
         winfunc=lambda n: np.ones((n,))):
    """base.py:8-16."""
    feat, energy = fbank(signal, samplerate, winlen, winstep, nfilt, nfft, lowfreq, highfreq, preemph, winfunc)
    feat = np.log(feat)
    feat = dct2_ortho(feat, numcep)
    feat = lifter(feat, ceplifter)
    if appendEnergy:
        feat[:, 0] = np.log(energy)
    return feat


def delta(feat, N):
    """base.py:70-79: regression over +-N edge-replicated frames."""
    if N < 1:
        raise ValueError("N must be an integer >= 1")
    feat = np.asarray(feat)
    nf = len(feat)
    denom = 2 * sum(i ** 2 for i in range(1, N + 1))
    padded = np.pad(feat, ((N, N), (0, 0)), mode="edge")
    out = np.zeros(feat.shape, dtype=np.result_type(feat.dtype, np.float64))
    for n in range(-N, N + 1):
        out += n * padded[N + n:N + n + nf]
    return (out / denom).astype(feat.dtype if feat.dtype.kind == "f" else np.float64)


def mfcc_delta39(signal, N=2, **kw):
    """[F,39] = mfcc | delta(mfcc,N) | delta(delta(mfcc,N),N): the fused kernel's contract
    (model.py:74-77 without the caller-side mean/scale)."""
    m = mfcc(signal, **kw)
    d1 = delta(m, N)
    d2 = delta(d1, N)
    return np.concatenate([m, d1, d2], axis=1)


# --------------------------------------------------------------------------
# preprocess.py
# --------------------------------------------------------------------------
def downsample_indices(n, src_rate, dst_rate):
    """Indices kept by `downsampling`, preprocess.py:21-28 (sample picking, no filter)."""
    f = np.arange(n, dtype=np.int64) * int(dst_rate) / int(src_rate)  # same int*int/int as the reference
    if dst_rate <= src_rate:
        # the k-th kept sample (k>=1) is the first i with f(i) > (k-1) + 1e-8; sample 0 always kept
        passed = np.where(f > 1e-8, np.floor(f - 1e-8) + 1, 0).astype(np.int64)
        # guard the floor against f-1e-8 landing exactly on an integer from below
        passed = np.where(f > passed + 1e-8, passed + 1, passed)
        passed = np.where(f > (passed - 1) + 1e-8, passed, passed - 1)
        keep = np.ones(n, dtype=bool)
        keep[1:] = passed[1:] > passed[:-1]
        return np.nonzero(keep)[0]
    cnt = -1
    out = []
    for i in range(n):
        if f[i] > cnt + 1e-8:
            cnt += 1
            out.append(i)
    return np.asarray(out, dtype=np.int64)


def downsampling(sig, src_rate, dst_rate):
    """preprocess.py:21-28."""
    sig = np.asarray(sig)
    return sig[downsample_indices(len(sig), src_rate, dst_rate)]


# --------------------------------------------------------------------------
# endpoint.py
# --------------------------------------------------------------------------
def get_amplitude(frames, window="square", use_sq=False):
    """endpoint.py:109-126.  Returns a Python list of np.float64 (Appendix A-13)."""
    frames = np.asarray(frames)
    l = frames[0].shape[-1]
    a = np.square(frames) if use_sq else np.abs(frames)
    if isinstance(window, str) and window == "hamming":
        w = np.hamming(l)
        a = np.stack([np.convolve(r, w, "same") for r in a])
    # 'square' convolves with ones(1): identity
    return [np.mean(r) for r in a]


def amplitude_feature(sig, rate, winlen, step):
    """endpoint.py:128-131."""
    return get_amplitude(to_frames(sig, rate, winlen, step))


def get_zcr(frames):
    """endpoint.py:182-198: strict sign changes between in-frame neighbours."""
    frames = np.asarray(frames)
    c = (frames[:, :-1] * frames[:, 1:] < 0).sum(axis=1)
    return [np.int64(v) for v in c]


def amplitude_rule(amp, mh=0.25, th=0.100, l_sil=0.100, r_sil=0.100, sigma=3, use_acr=False,
                   frames=None, rate=None, cfg_frame=CFG_FRAME, cfg_step=CFG_STEP):
    """endpoint.py:133-179: double-threshold state machine, returns list of (j,k)."""
    amp = list(amp)

    def acr_rule(frame):
        acrs = [acr(frame, n) for n in range(rate // 500, rate // 50)]
        return max(acrs) / acr(frame, 0) > 0.55

    p = []
    sil = amp[:int(l_sil / cfg_step)] + amp[-int(r_sil / cfg_step):]
    sil = sorted(sil)[:-2]
    s_mean, s_sigma = np.mean(sil), np.std(sil)
    T_H = th / cfg_frame
    M_L = s_mean + sigma * s_sigma
    M_H = max(np.max(amp) * mh, M_L)
    i = 0
    n = len(amp)
    while i < n:
        if amp[i] >= M_H:
            j = k = i
            while k < n and amp[k] > M_H:
                k += 1
            if k - j < T_H:
                i = k
            else:
                while j > 0 and amp[j] > M_L and (not use_acr or acr_rule(frames[j])):
                    j -= 1
                while k < n and amp[k] > M_L and (not use_acr or acr_rule(frames[k])):
                    k += 1
                p.append((j, k))
                i = k
        i += 1
    if not p:
        return [(0, n)]
    return p


def zcr_rule(zcr, left, right, max_shift=0.400, l_sil=0, r_sil=0.100, cfg_frame=CFG_FRAME, cfg_step=CFG_STEP):
    """endpoint.py:201-220."""
    zcr = list(zcr)
    max_shift /= cfg_frame
    sil = zcr[:int(l_sil / cfg_step)] + zcr[-int(r_sil / cfg_step):]
    mu, sg = np.mean(sil), np.std(sil)
    thres = mu + 3 * sg
    j = left
    while j > 0 and left - j <= max_shift and zcr[j] > thres:
        j -= 1
    k = right
    while k < len(zcr) and k - right <= max_shift and zcr[k] > thres:
        k += 1
    return j, k


def basic_endpoint_detection(sig, rate, return_feature=False, cfg_frame=CFG_FRAME, cfg_step=CFG_STEP):
    """endpoint.py:34-66."""
    frames = to_frames(sig, rate, t=cfg_frame, step=cfg_step)
    amp = get_amplitude(frames)
    kw = dict(cfg_frame=cfg_frame, cfg_step=cfg_step)
    sep = amplitude_rule(amp, **kw)
    left, right = sep[0][0], sep[-1][1]
    if right - left < 50:
        sep = amplitude_rule(amp, 0.125, **kw)
    left, right = sep[0][0], sep[-1][1]
    zcr = get_zcr(frames)
    left2, right2 = zcr_rule(zcr, left, right, **kw)
    if right2 - left2 < 50:
        left2, right2 = 0, len(frames)
    res = int(left2 * cfg_step * rate), int(right2 * cfg_step * rate)
    return res if not return_feature else res + (amp, zcr)


def robust_endpoint_detection(sig, rate, cfg_frame=CFG_FRAME, cfg_step=CFG_STEP):
    """endpoint.py:68-92 (autocorrelation-gated expansion, mh=0.5)."""
    frames = to_frames(sig, rate, cfg_frame, step=cfg_step)
    amp = get_amplitude(frames)
    kw = dict(cfg_frame=cfg_frame, cfg_step=cfg_step)
    sep = amplitude_rule(amp, 0.5, frames=frames, use_acr=True, rate=rate, **kw)
    left, right = sep[0][0], sep[-1][1]
    zcr = get_zcr(frames)
    left2, right2 = zcr_rule(zcr, left, right, **kw)
    if right2 - left2 < 50:
        left2, right2 = 0, len(frames)
    return int(left2 * cfg_step * rate), int(right2 * cfg_step * rate)


def get_noise(amp, sep_point):
    """endpoint.py:94-107: mean amplitude outside the detected segments (1e30 for the whole-signal fallback)."""
    n = len(amp)
    if sep_point[0] == (0, n):
        return 1e30
    gaps = []                                   # (start, stop) of every stretch between / around the segments
    cursor = 0
    for seg_start, seg_stop in sep_point:
        gaps.append((cursor, seg_start))
        cursor = seg_stop
    gaps.append((cursor, n))
    acc, frames = 0, 0
    for g0, g1 in gaps:                         # same accumulation order as the reference's running sum
        acc += np.sum(amp[g0:g1])
        frames += g1 - g0
    return acc / frames


def endpoint_max_pitch(l, rate, bias=20):
    """endpoint.py:15-18."""
    idx = bias + np.argmax(l)
    return 1 / (1.0 / rate * idx)


# --------------------------------------------------------------------------
# pitch.py
# --------------------------------------------------------------------------
def center_clip(frame, binary=True):
    """pitch.py:145-155 (== endpoint.py:20-30).  Accepts [N] or [F,N]."""
    frame = np.asarray(frame)
    one = frame.ndim == 1
    fr = np.atleast_2d(frame).astype(np.float64)
    out = np.zeros(fr.shape, dtype=np.float64)
    for r in range(fr.shape[0]):
        x = fr[r]
        with np.errstate(all="ignore"):
            med = np.median(x[x >= 0]) if np.any(x >= 0) else np.nan
        hi = x > med
        lo = (~hi) & (x < -med)
        if binary:
            out[r][hi] = 1
            out[r][lo] = -1
        else:
            out[r][hi] = x[hi] - med
            out[r][lo] = x[lo] + med
    if binary:
        out = out.astype(np.int64)
    return out[0] if one else out


def pitch_detect_frame(frame, rate, gender="male"):
    """pitch.py:135-143: |ifft(log|fft(window(frame,50,1000,'hamming'))|)|.  [N] or [F,N]."""
    y = window(frame, rate, 50, 1000, "hamming")
    with np.errstate(divide="ignore", invalid="ignore"):
        log_Xw = np.log(np.abs(np.fft.fft(y, axis=-1)))
        return np.abs(np.fft.ifft(log_Xw, axis=-1))


def pitch_detect_frame_sr(frame, rate, min_shift=20, max_shift=200):
    """pitch.py:112-132: unbiased autocorrelation of |band-passed frame| at lags 20..199."""
    v = np.abs(window(frame, rate, 50, 900, "hamming"))
    one = v.ndim == 1
    v = np.atleast_2d(v)
    L = v.shape[-1]
    sc = np.stack([np.sum(v[:, :-n] * v[:, n:], axis=1) / (L - n) for n in range(min_shift, max_shift)], axis=1)
    return list(sc[0]) if one else sc


def smooth(g, degree=2):
    """pitch.py:157-164: in-place running mean => recurrence over already-smoothed rows."""
    g = np.array(g, dtype=np.float64)
    L = len(g)
    for i in range(L):
        left = i - degree if i - degree >= 0 else 0
        right = i + degree if i + degree < L else L - 1
        with np.errstate(all="ignore"):
            g[i] = np.mean(g[left:right], axis=0)
    return g.tolist()


def peak_score(sig, gender="male", min_f=20, max_f=100):
    """pitch.py:227-242: distance to the nearest strictly greater sample on either side."""
    sig = np.asarray(sig, dtype=np.float64)
    n = len(sig)
    out = []
    for i in range(min_f, max_f):
        v = sig[i]
        with np.errstate(invalid="ignore"):
            stop = ~(sig <= v)  # NaN compares False => stops immediately (Appendix A-11)
        left = np.nonzero(stop[1:i + 1])[0]
        p = (left[-1] + 1) if len(left) else 0
        right = np.nonzero(stop[i:])[0]
        q = (right[0] + i) if len(right) else n
        out.append(int(min(i - p, q - i)))
    return out


def max_pitch(g, bias=20):
    """pitch.py:166-172."""
    return [1 / (0.0001 * (bias + np.argmax(l))) for l in g]


def greedy_max_pitch(g, bias=20):
    """pitch.py:174-189."""
    pitch = []
    for l in g:
        p = 0
        for i in range(len(l) - 1):
            if l[i] > l[i + 1]:
                p = 1 / (0.0001 * (i + bias))
                break
        pitch.append(p)
    return pitch


def robust_max_pitch(g, bias=20):
    """pitch.py:191-206: octave-error repair, forward then backward sweep."""
    C = 50
    pitch = max_pitch(g, bias)
    for i in range(1, len(pitch)):
        if abs(2 * pitch[i] - pitch[i - 1]) < C and pitch[i] < 170:
            pitch[i] = 2 * pitch[i]
    for i in range(len(pitch) - 2, 0, -1):
        if abs(2 * pitch[i] - pitch[i + 1]) < C and pitch[i] < 170:
            pitch[i] = 2 * pitch[i]
    return pitch


def robust_pitch_from_lags(lags):
    """robust_max_pitch (pitch.py:191-206) restarted from integer lags (= bias + argmax): Hz per frame."""
    pitch = [1 / (0.0001 * int(l)) for l in lags]
    for i in range(1, len(pitch)):
        if abs(2 * pitch[i] - pitch[i - 1]) < 50 and pitch[i] < 170:
            pitch[i] = 2 * pitch[i]
    for i in range(len(pitch) - 2, 0, -1):
        if abs(2 * pitch[i] - pitch[i + 1]) < 50 and pitch[i] < 170:
            pitch[i] = 2 * pitch[i]
    return pitch


def dp_max_pitch(g):
    """pitch.py:208-225 (Viterbi over lags; never called on the path)."""
    g = np.array(g)
    dp = np.zeros(g.shape)
    prev = np.zeros(g.shape, dtype=int)
    k = np.arange(g.shape[1])
    step = 0
    for i in range(1, dp.shape[0]):
        for j in range(dp.shape[1]):
            reward = dp[i - 1] - 5 * np.abs(k - j) + g[i][j]
            step = int(np.argmax(reward))
            dp[i][j] = reward[step]
            prev[i][j] = step
    i = dp.shape[0] - 1
    path = np.zeros(g.shape[0])
    while i >= 0:
        with np.errstate(divide="ignore"):
            path[i] = np.float64(10000) / step
        step = prev[i][step]
        i -= 1
    return path.tolist()


def pitch_scores_cep(sig, rate, winlen=0.0512, step=0.01):
    """pitch_detect up to (not including) robust_max_pitch: returns (scores [F,80] int, frames)."""
    sig = downsampling(sig, rate, 10000)
    frames = to_frames(sig, 10000, winlen, step)
    clipped = center_clip(frames, False)
    ceps = pitch_detect_frame(clipped, 10000)
    ceps = np.asarray(smooth(ceps))
    scores = [peak_score(c) for c in ceps]
    return scores, frames


def pitch_detect(sig, rate, winlen=0.0512, step=0.01, gender="male"):
    """pitch.py:83-94: cepstrum pitch.  Returns (list of Hz, frames)."""
    scores, frames = pitch_scores_cep(sig, rate, winlen, step)
    return robust_max_pitch(scores), frames


def pitch_scores_sr(sig, rate, winlen=0.0512, step=0.01):
    """pitch_detect_sr up to robust_max_pitch: returns (smoothed scores [F,180], frames)."""
    sig = downsampling(sig, rate, 10000)
    frames = to_frames(sig, 10000, winlen, step)
    clipped = center_clip(frames, False)
    sc = pitch_detect_frame_sr(clipped, 10000)
    return smooth(sc, 2), frames


def pitch_detect_sr(sig, rate, winlen=0.0512, step=0.01):
    """pitch.py:96-110: autocorrelation pitch."""
    scores, frames = pitch_scores_sr(sig, rate, winlen, step)
    return robust_max_pitch(scores, bias=20), frames


def pitch_rows_cep(sig, rate, winlen=0.0512, step=0.01):
    """The smoothed cepstrum rows peak_score sees inside pitch_detect (pitch.py:83-91): float64 [F, N]."""
    sig = downsampling(sig, rate, 10000)
    frames = to_frames(sig, 10000, winlen, step)
    return np.asarray(smooth(pitch_detect_frame(center_clip(frames, False), 10000)))


def peak_score_bounds(row, tol, min_f=20, max_f=100):
    """peak_score (pitch.py:227-242) with every comparison `row[j] <= v` moved by -tol / +tol: (lo, hi) integer score bounds
    that any evaluation of the row perturbed by less than tol / 2 per sample must respect.  Checker for float32 near-ties."""
    row = np.asarray(row, dtype=np.float64)
    n = len(row)
    lo, hi = [], []
    for i in range(min_f, max_f):
        v = row[i]
        out = []
        for t in (-tol, tol):
            with np.errstate(invalid="ignore"):
                stop = ~(row <= v + t)
            stop[i] = False
            left = np.nonzero(stop[1:i + 1])[0]
            p = (left[-1] + 1) if len(left) else 0
            right = np.nonzero(stop[i:])[0]
            q = (right[0] + i) if len(right) else n
            out.append(int(min(i - p, q - i)))
        lo.append(out[0]); hi.append(out[1])
    return np.array(lo), np.array(hi)


def lag_is_near_tie_cep(row, lag, rel_tol=1e-5, bias=20):
    """True when `lag` (= bias + argmax of some float32 evaluation of peak_score(row)) is explained by a near-tie of the
    float64 row: its upper score bound reaches the largest lower bound (north_star: mismatches are allowed only where the
    reference statistic lies within tolerance of the decision threshold)."""
    row = np.asarray(row, dtype=np.float64)
    if not np.all(np.isfinite(row)):
        return False
    lo, hi = peak_score_bounds(row, rel_tol * float(np.max(np.abs(row))))
    return bool(hi[int(lag) - bias] >= lo.max())


def lag_is_near_tie_sr(row, lag, rel_tol=1e-5, bias=20):
    """Autocorrelation rows (pitch.py:96-110): the chosen lag's smoothed score lies within rel_tol * max|row| of the maximum."""
    row = np.asarray(row, dtype=np.float64)
    if not np.all(np.isfinite(row)):
        return False
    return bool(row[int(lag) - bias] >= row.max() - rel_tol * float(np.max(np.abs(row))))


def sub_endpoint_detect(frames):
    """pitch.py:64-81: deepest +-10-frame amplitude valley."""
    amp = np.array([np.abs(f).sum() for f in frames])
    p, max_score = 0, -1000
    for i in range(10, len(amp) - 10):
        if np.any(amp[i - 2:i + 3] < amp[i]):
            continue
        s = sum(amp[j] - amp[i] for j in range(i - 10, i + 11))
        if s > max_score:
            max_score, p = s, i
    return len(amp) // 2 if p == 0 else p


def find_smooth_subsequence(pitch, base_tor=3, base_thres=30, bias=0):
    """pitch.py:245-279: longest run with at most `tor` jumps > `thres` Hz."""
    pitch = list(pitch)
    tor, thres = base_tor, base_thres
    i, strs, idx = 0, [], []
    while i < len(pitch):
        j, prev, k, seg = i + 1, pitch[i], tor, [pitch[i]]
        while j < len(pitch):
            if abs(pitch[j] - prev) > thres:
                k -= 1
            else:
                seg.append(pitch[j])
                prev = pitch[j]
            if not k:
                strs.append(seg)
                idx.append((i + bias, j + bias))
                break
            j += 1
        if j == len(pitch):
            strs.append(seg)
            idx.append((i + bias, j + bias))
            break
        i = j - tor + 1
    obj = sorted(zip(strs, idx), key=lambda x: -len(x[0]))
    strs, idx = tuple(zip(*obj))
    return strs[0], idx[0]


def slope(seq):
    """pitch.py:49-52."""
    return np.polyfit(np.arange(0, len(seq)), seq, 1)[0]


def quad_params(seq):
    """pitch.py:54-57."""
    return np.polyfit(np.arange(0, len(seq)), seq, 2)[0]


def peakshift(seq1, seq2):
    """pitch.py:59-62."""
    return np.median(seq2) - np.median(seq1)


def pitch_feature(sig, rate, gender="male"):
    """pitch.py:26-47 without the two stdout prints.  Returns the 5-tuple SVM input."""
    pitch, frames = pitch_detect(sig, rate)
    p = sub_endpoint_detect(frames)
    p_bias = 5 if p > 15 else 0
    s1, _ = find_smooth_subsequence(pitch[p_bias:p], bias=p_bias)
    s2, _ = find_smooth_subsequence(pitch[p:], bias=p)
    return slope(s1), slope(s2), quad_params(s1), quad_params(s2), peakshift(s1, s2)


# --------------------------------------------------------------------------
# model.py glue around the features (the "next" row f-1 of SURVEY.md section 8)
# --------------------------------------------------------------------------
def sk_scale(x, with_mean=True):
    """sklearn.preprocessing.scale(x) on a 2-D array, axis 0: population std, constant columns are not divided."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(axis=0) if with_mean else np.zeros(x.shape[1])
    sd = x.std(axis=0)
    sd = np.where(sd < 10 * np.finfo(np.float64).eps, 1.0, sd)
    return (x - mean) / sd


def model_batch(utterances, rate=16000, T=200, N=3, scale_signal=True, **mfcc_kw):
    """model.py:52-64 (endpoint_detect without augmentation), :66-88 (feature_extract_mfcc), :35-50 (pad), :114-135
    (get_batch_full): returns (inp float64 [T, B, 39], len0 int array)."""
    feats, len0 = [], []
    for sig in utterances:
        l, r = basic_endpoint_detection(sig, rate)
        sound = np.asarray(sig[l:r], dtype=np.float64).reshape(-1, 1)
        if scale_signal:
            sound = sk_scale(sound, with_mean=False)              # model.py:62-63
        m0 = mfcc(sound.reshape(-1), rate, **mfcc_kw)
        m0 = m0 - np.mean(m0)                                     # model.py:75
        d1 = delta(m0, N)
        d2 = delta(d1, N)
        m0 = sk_scale(m0)                                         # model.py:78
        f = np.concatenate([m0, d1, d2], axis=1)
        f = np.pad(f, ((0, T - len(f)), (0, 0))) if len(f) < T else f[:T]   # model.py:35-39
        feats.append(f)
        len0.append(min(len(m0), T))
    return np.stack(feats).transpose(1, 0, 2), np.array(len0)
