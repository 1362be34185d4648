"""Shared plumbing of the drop-in package: plan caches, dtype routing, reference-side config lookup."""
import functools

import numpy as np

import dspfe


def cfg_frame_step():
    """cfg.frame / cfg.step of the caller's `config` module (reference config.py:31-32; read inside endpoint.py),
    or the reference defaults when no such module is importable."""
    try:
        from config import cfg  # the host project's global config, if present
        return float(cfg.frame), float(cfg.step)
    except Exception:
        return 0.03, 0.01


def pack_one(sig):
    """One utterance as a packed batch of one.  int16-valued input keeps the int16 kernels; anything else goes
    through the float32 sample path.  Returns (array, is_f32)."""
    a = np.asarray(sig)
    if a.ndim != 1:
        raise NotImplementedError("expected a 1-D signal")
    if a.dtype == np.int16:
        return np.ascontiguousarray(a), False
    if a.dtype.kind in "iu" and (a.size == 0 or (a.min() >= -32768 and a.max() <= 32767)):
        return a.astype(np.int16), False
    if a.dtype.kind == "f":
        if a.size and np.all(np.abs(a) <= 32767) and np.array_equal(a, np.rint(a)):
            return a.astype(np.int16), False
        return a.astype(np.float32), True
    raise NotImplementedError(f"unsupported signal dtype {a.dtype}")


@functools.lru_cache(maxsize=32)
def _mfcc_plan(key):
    kw = dict(key)
    win = kw.pop("window")
    return dspfe.MfccPlan(window=None if win is None else np.frombuffer(win, dtype=np.float64), **kw)


def mfcc_plan(window=None, **kw):
    w = None
    if window is not None:
        window = np.ascontiguousarray(window, dtype=np.float64)
        if not np.all(window == 1.0):
            w = window.tobytes()
    return _mfcc_plan(tuple(sorted(dict(kw, window=w).items())))


@functools.lru_cache(maxsize=8)
def endpoint_plan(rate, cfg_frame, cfg_step):
    return dspfe.EndpointPlan(samplerate=int(rate), cfg_frame=cfg_frame, cfg_step=cfg_step)


def to_device(arr):
    import torch
    if not torch.cuda.is_available():
        raise dspfe.DspfeError(-3, "no CUDA device: the features package has no CPU fallback")
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()
