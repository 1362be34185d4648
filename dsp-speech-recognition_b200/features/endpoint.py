"""Drop-in for reference features/endpoint.py: energy + zero-crossing endpoint detection."""
import numpy as np

import dspfe
from . import _gpu
from .sigproc import *  # noqa: F401,F403  (the reference module star-imports sigproc, endpoint.py:8)
from .preprocess import preemphasis  # noqa: F401

try:  # names the reference module re-exports from the host project (endpoint.py:9-13); optional here
    from plotter import plot_frame, show  # noqa: F401
except Exception:
    def plot_frame(*a, **k):
        return None

    def show(*a, **k):
        return None
try:
    from reader import Reader  # noqa: F401
except Exception:
    Reader = None
try:
    from config import cfg, meta  # noqa: F401
except Exception:
    cfg = meta = None

noise_rec = []


def max_pitch(l, rate, bias=20):
    """reference endpoint.py:15-18."""
    idx = bias + np.argmax(l)
    return 1 / (1.0 / rate * idx)


def center_clip(frame, binary=True):
    """reference endpoint.py:20-30 (shadowed by pitch.center_clip in the flat namespace)."""
    from .pitch import center_clip as _cc
    return _cc(frame, binary)


def basic_endpoint_detection(sig, rate, return_feature=False):
    """reference endpoint.py:34-66 on the device (K2a/K2b/K3).  Returns Python ints (left, right) in samples, plus
    the amplitude list (np.float64) and zero-crossing list (np.int64) when return_feature."""
    cfg_frame, cfg_step = _gpu.cfg_frame_step()
    x, f32 = _gpu.pack_one(sig)
    if f32:
        return _float_endpoint_detection(np.asarray(sig, dtype=np.float64), rate, cfg_frame, cfg_step, return_feature, False)
    plan = _gpu.endpoint_plan(int(rate), cfg_frame, cfg_step)
    off = np.array([0, len(x)], dtype=np.int64)
    if not return_feature:
        lr = plan.detect_host(x, off)
        return int(lr[0, 0]), int(lr[0, 1])
    lr, asum, zcr, _ = plan.detect_host(x, off, want_features=True)
    amp = [np.float64(v) / np.float64(plan.frame_len) for v in asum]
    return int(lr[0, 0]), int(lr[0, 1]), amp, [np.int64(v) for v in zcr]


def robust_endpoint_detection(sig, rate):
    """reference endpoint.py:68-92 on the device: amplitude_rule(mh=0.5) gated by the autocorrelation peak of each frame
    (exact integer lag sums), zcr_rule, whole-signal fallback.  Returns Python ints (left, right) in samples."""
    cfg_frame, cfg_step = _gpu.cfg_frame_step()
    x, f32 = _gpu.pack_one(sig)
    if f32:
        return _float_endpoint_detection(np.asarray(sig, dtype=np.float64), rate, cfg_frame, cfg_step, False, True)
    plan = _gpu.endpoint_plan(int(rate), cfg_frame, cfg_step)
    lr = plan.detect_robust_host(x, np.array([0, len(x)], dtype=np.int64))
    return int(lr[0, 0]), int(lr[0, 1])


def _float_endpoint_detection(sig, rate, cfg_frame, cfg_step, return_feature, robust):
    """Signals that are not int16-valued (reference endpoint.py:34 takes any real dtype): the reference's own sequence of
    calls (endpoint.py:40-66 / :71-92) over this module's device-backed pieces -- framing, per-frame amplitude and
    zero-crossing kernels in float64, then the shared C++ rules.  The fused int16 kernels K2/K3 are the batch path."""
    frames = to_frames(sig, rate, t=cfg_frame, step=cfg_step)  # noqa: F405
    amp = get_amplitude(frames)
    if robust:
        sep_point = amplitude_rule(amp, 0.5, frames=frames, use_acr=True, rate=rate)
    else:
        sep_point = amplitude_rule(amp)
        left, right = sep_point[0][0], sep_point[-1][1]
        if right - left < 50:
            sep_point = amplitude_rule(amp, 0.125)
    left, right = sep_point[0][0], sep_point[-1][1]
    zcr = get_zcr(frames)
    left2, right2 = zcr_rule(zcr, left, right)
    if right2 - left2 < 50:
        left2, right2 = 0, len(frames)
    lr = int(left2 * cfg_step * rate), int(right2 * cfg_step * rate)
    return lr + (amp, zcr) if return_feature else lr


def get_noise(amp, sep_point):
    """reference endpoint.py:94-107: mean amplitude of the frames outside every detected (j, k) segment; 1e30 when the
    rule fell back to the whole signal."""
    amp = np.asarray(amp, dtype=np.float64)
    n = len(amp)
    if tuple(sep_point[0]) == (0, n):
        return 1e30
    # the reference walks the gaps between consecutive segments; accumulate them in the same order
    total, count, start = 0, 0, 0
    for j, k in sep_point:
        total = total + np.sum(amp[start:j])
        count += j - start
        start = k
    return (total + np.sum(amp[start:])) / (count + n - start)


def get_amplitude(frames, window='square', use_sq=False):
    """reference endpoint.py:109-126 on the device: per frame the mean of |x| (or x^2), optionally convolved ('same')
    with a Hamming window first.  Returns a list of np.float64."""
    frames = np.asarray(frames, dtype=np.float64)
    if isinstance(window, str) and window == 'square':
        return [np.float64(v) for v in dspfe.row_amplitude_f64(frames, use_sq)]
    if isinstance(window, str) and window == 'hamming':
        return [np.float64(v) for v in dspfe.row_windowed_amplitude_f64(frames, np.hamming(frames.shape[-1]), use_sq)]
    raise NotImplementedError("window must be 'square' or 'hamming'")


def amplitude_feature(sig, rate, winlen, step):
    """reference endpoint.py:128-131."""
    return get_amplitude(to_frames(sig, rate, winlen, step))  # noqa: F405


def amplitude_rule(amp, mh=0.25, th=0.100, l_sil=0.100, r_sil=0.100, sigma=3, use_acr=False, frames=None, rate=None):
    """reference endpoint.py:133-179: list of (j, k) frame segments, or [(0, len(amp))].  Runs the C++ rule that
    the device kernel K3 runs (csrc/endpoint_kernel.cuh), through dspfe_amplitude_rule_host."""
    cfg_frame, cfg_step = _gpu.cfg_frame_step()
    if use_acr:   # acr_rule of every frame on the device, then the same C++ rule with the gate
        gate = dspfe.acr_gate_rows_f64(np.asarray(frames, dtype=np.float64), int(rate))
        return dspfe.amplitude_rule_gated_host([float(a) for a in amp], gate, mh, cfg_frame=cfg_frame, cfg_step=cfg_step, th=float(th),
                                               l_sil=float(l_sil), r_sil=float(r_sil), sigma=float(sigma))
    return dspfe.amplitude_rule_host([float(a) for a in amp], mh, cfg_frame=cfg_frame, cfg_step=cfg_step, th=float(th),
                                     l_sil=float(l_sil), r_sil=float(r_sil), sigma=float(sigma))


def get_zcr(frames):
    """reference endpoint.py:182-198 on the device (dspfe_row_zcr_f64).  Returns a list of np.int64."""
    frames = np.asarray(frames, dtype=np.float64)
    return [np.int64(v) for v in dspfe.row_zcr_f64(frames)]


def zcr_rule(zcr, left, right, max_shift=0.400, l_sil=0, r_sil=0.100):
    """reference endpoint.py:201-220 (same C++ rule as K3, through dspfe_zcr_rule_host)."""
    cfg_frame, cfg_step = _gpu.cfg_frame_step()
    return dspfe.zcr_rule_host([float(z) for z in zcr], int(left), int(right), l_sil=float(l_sil), cfg_frame=cfg_frame,
                               cfg_step=cfg_step, zcr_max_shift=float(max_shift), zcr_r_sil=float(r_sil))
