"""Drop-in for reference features/sigproc.py (framing, power spectra, pre-emphasis, FIR band-pass, autocorrelation)."""
import decimal
import logging
import math

import numpy
import numpy as np

import dspfe
from . import _gpu


def to_frames(sig, rate, t=0.020, step=0.010):
    """reference sigproc.py:11-19: frames of int(rate*t) samples every int(step*rate)."""
    return framesig(sig, int(rate * t), int(step * rate))


def window(sig, rate, low_freq=0, high_freq=500, wintype='square'):
    """reference sigproc.py:22-46: causal complex FIR band-pass (one-sided ideal band => complex taps), output truncated
    to len(sig); float64 on the device.  The fused pitch kernels apply the same filter through its 1024-point spectrum."""
    sig = np.asarray(sig)
    if sig.ndim != 1:
        raise NotImplementedError("expected a 1-D signal")
    return dspfe.fir_window_f64(sig, rate, low_freq, high_freq, wintype == 'hamming')


def acr(frame, n):
    """reference sigproc.py:48-53: unbiased autocorrelation at lag n (n = 0: mean square)."""
    return dspfe.acr_f64(np.asarray(frame, dtype=np.float64), n)


def round_half_up(number):
    """reference sigproc.py:55-56."""
    return int(decimal.Decimal(number).quantize(decimal.Decimal('1'), rounding=decimal.ROUND_HALF_UP))


def rolling_window(a, window, step=1):
    """reference sigproc.py:59-63 (strided view)."""
    shape = a.shape[:-1] + (a.shape[-1] - window + 1, window)
    strides = a.strides + (a.strides[-1],)
    return numpy.lib.stride_tricks.as_strided(a, shape=shape, strides=strides)[::step]


def framesig(sig, frame_len, frame_step, winfunc=lambda x: numpy.ones((x,)), stride_trick=True):
    """reference sigproc.py:66-98, on the device (dspfe_frames_f64).  Returns float64 [NUMFRAMES, frame_len]."""
    frame_len = int(round_half_up(frame_len))
    frame_step = int(round_half_up(frame_step))
    win = numpy.asarray(winfunc(frame_len), dtype=numpy.float64)
    return dspfe.frames_f64(numpy.asarray(sig).astype(numpy.float64), frame_len, frame_step,
                            None if numpy.all(win == 1.0) else win)


def deframesig(frames, siglen, frame_len, frame_step, winfunc=lambda x: numpy.ones((x,))):
    """reference sigproc.py:101-133: overlap-add inverse of framesig (host; not on the feature path)."""
    frame_len = round_half_up(frame_len)
    frame_step = round_half_up(frame_step)
    numframes = numpy.shape(frames)[0]
    assert numpy.shape(frames)[1] == frame_len, '"frames" matrix is wrong size, 2nd dim is not equal to frame_len'
    padlen = (numframes - 1) * frame_step + frame_len
    if siglen <= 0:
        siglen = padlen
    rec_signal = numpy.zeros((padlen,))
    window_correction = numpy.zeros((padlen,))
    win = winfunc(frame_len)
    for i in range(numframes):
        sl = slice(i * frame_step, i * frame_step + frame_len)
        window_correction[sl] += win + 1e-15
        rec_signal[sl] += frames[i, :]
    return (rec_signal / window_correction)[0:siglen]


def _spectrum(frames, NFFT, kind):
    frames = numpy.asarray(frames, dtype=numpy.float64)
    if frames.ndim != 2:
        raise NotImplementedError("frames must be a 2-D array")
    if NFFT not in (32, 64, 128, 256, 512, 1024, 1536, 2048):
        raise NotImplementedError("NFFT must be a power of two in [32, 2048] or 1536")
    if numpy.shape(frames)[1] > NFFT:
        logging.warning('frame length (%d) is greater than FFT size (%d), frame will be truncated. Increase NFFT to avoid.',
                     numpy.shape(frames)[1], NFFT)
        frames = frames[:, :NFFT]
    nf, L = frames.shape
    L2 = L + (L & 1)                      # the kernel wants an even hop; an extra zero sample changes nothing
    buf = numpy.zeros((nf, L2), dtype=numpy.float32)
    buf[:, :L] = frames
    plan = _gpu.mfcc_plan(frame_len=L2, frame_step=L2, preemph=0.0, nfft=NFFT)
    off = _gpu.to_device(numpy.arange(nf + 1, dtype=numpy.int64) * L2)
    out, _ = plan.spectrum_f32(_gpu.to_device(buf.reshape(-1)), off, kind)
    return out[:nf].cpu().numpy().astype(numpy.float64)


def magspec(frames, NFFT):
    """reference sigproc.py:136-148: |rfft(frames, NFFT)|."""
    return _spectrum(frames, NFFT, 1)


def powspec(frames, NFFT):
    """reference sigproc.py:151-158: |rfft|^2 / NFFT."""
    return _spectrum(frames, NFFT, 0)


def logpowspec(frames, NFFT, norm=1):
    """reference sigproc.py:161-175: 10*log10(max(ps, 1e-30)), minus the global maximum when norm."""
    lps = _spectrum(frames, NFFT, 2)
    return lps - numpy.max(lps) if norm else lps


def preemphasis(signal, coeff=0.95):
    """reference sigproc.py:178-185.  A 2-D (1,S) input comes back flattened and unfiltered, as in the reference."""
    signal = numpy.asarray(signal)
    if signal.ndim == 2 and signal.shape[0] == 1:
        return signal[0].copy()
    if signal.ndim != 1:
        raise NotImplementedError("expected a 1-D signal (or the (1,S) form the reference callers use)")
    return dspfe.preemphasis_f64(signal.astype(numpy.float64), coeff)
