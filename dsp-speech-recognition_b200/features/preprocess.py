"""Drop-in for reference features/preprocess.py."""
import numpy as np

import dspfe


def preemphasis(signal, coeff=0.95):
    """reference preprocess.py:11-19 (same filter as sigproc.preemphasis)."""
    signal = np.asarray(signal)
    if signal.ndim == 2 and signal.shape[0] == 1:
        return signal[0].copy()
    if signal.ndim != 1:
        raise NotImplementedError("expected a 1-D signal")
    return dspfe.preemphasis_f64(signal.astype(np.float64), coeff)


def downsample_indices(n, src_rate, dst_rate):
    """Indices kept by `downsampling` (sample picking, no anti-alias filter): sample i survives iff
    i*dst/src > cnt + 1e-8 with cnt the number kept so far minus one (reference preprocess.py:21-28)."""
    f = np.arange(n, dtype=np.int64) * int(dst_rate) / int(src_rate)
    keep = np.zeros(n, dtype=bool)
    cnt = -1
    if dst_rate <= src_rate:
        passed = np.where(f > 1e-8, np.floor(f - 1e-8) + 1, 0).astype(np.int64)
        passed = np.where(f > passed + 1e-8, passed + 1, passed)
        passed = np.where(f > (passed - 1) + 1e-8, passed, passed - 1)
        keep[:1] = n > 0
        keep[1:] = passed[1:] > passed[:-1]
        return np.nonzero(keep)[0]
    for i in range(n):
        if f[i] > cnt + 1e-8:
            cnt += 1
            keep[i] = True
    return np.nonzero(keep)[0]


def downsampling(sig, src_rate, dst_rate):
    """reference preprocess.py:21-28.  Pure index selection (host); the pitch kernels fold the same index map
    into their sample gather."""
    sig = np.asarray(sig)
    return sig[downsample_indices(len(sig), src_rate, dst_rate)]
