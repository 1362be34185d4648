"""Drop-in for reference features/base.py: MFCC, filterbank energies, deltas."""
import logging

import numpy

import dspfe
from . import _gpu, sigproc
from .sigproc import to_frames  # noqa: F401  (re-exported by the reference module)

try:  # the reference namespace exposes scipy's dct; kept for star-import parity only
    from scipy.fftpack import dct  # noqa: F401
except Exception:  # pragma: no cover
    dct = None


_NFFT_BUILT = (32, 64, 128, 256, 512, 1024, 1536, 2048)


def _prepare(signal, samplerate, winlen, winstep, nfft, preemph, winfunc):
    signal = numpy.asarray(signal)
    if signal.ndim == 2:
        if signal.shape[0] != 1:
            raise NotImplementedError("2-D signals other than (1,S) are not supported")
        # reference sigproc.py:185 on a (1,S) array: signal[1:] is empty, so the row comes back unfiltered
        signal, preemph = signal[0], 0.0
    if nfft not in _NFFT_BUILT:
        raise NotImplementedError("nfft must be a power of two in [32, 2048] or 1536 (model.py:74)")
    frame_len = sigproc.round_half_up(winlen * samplerate)
    frame_step = sigproc.round_half_up(winstep * samplerate)
    if frame_len > nfft:   # reference sigproc.py:143-146: frames are counted and windowed at full length, the transform takes their first nfft samples
        logging.warning('frame length (%d) is greater than FFT size (%d), frame will be truncated. Increase NFFT to avoid.', frame_len, nfft)
    win = numpy.asarray(winfunc(frame_len), dtype=numpy.float64)
    x, f32 = _gpu.pack_one(signal)
    return x, f32, frame_len, frame_step, float(preemph), win


def mfcc(signal, samplerate=16000, winlen=0.025, winstep=0.01, numcep=13,
         nfilt=26, nfft=512, lowfreq=0, highfreq=None, preemph=0.97, ceplifter=22, appendEnergy=True,
         winfunc=lambda x: numpy.ones((x,))):
    """reference base.py:8-16.  Returns float64 [NUMFRAMES, numcep]."""
    highfreq = highfreq or samplerate / 2
    assert highfreq <= samplerate / 2, "highfreq is greater than samplerate/2"
    x, f32, flen, fstep, pre, win = _prepare(signal, samplerate, winlen, winstep, nfft, preemph, winfunc)
    plan = _gpu.mfcc_plan(samplerate=samplerate, frame_len=flen, frame_step=fstep, nfft=nfft, nfilt=nfilt, numcep=numcep,
                          ceplifter=int(ceplifter), append_energy=bool(appendEnergy), delta_n=1, preemph=pre,
                          lowfreq=float(lowfreq), highfreq=float(highfreq), window=win)
    off = _gpu.to_device(numpy.array([0, len(x)], dtype=numpy.int64))
    out, fo = (plan.mfcc_delta_f32 if f32 else plan.mfcc_delta)(_gpu.to_device(x), off)
    n = dspfe.num_frames(len(x), flen, fstep)
    return out[:n, :numcep].cpu().numpy().astype(numpy.float64)


def fbank(signal, samplerate=16000, winlen=0.025, winstep=0.01,
          nfilt=26, nfft=512, lowfreq=0, highfreq=None, preemph=0.97,
          winfunc=lambda x: numpy.ones((x,))):
    """reference base.py:18-32.  Returns (feat float64 [NUMFRAMES, nfilt], energy float64 [NUMFRAMES])."""
    highfreq = highfreq or samplerate / 2
    assert highfreq <= samplerate / 2, "highfreq is greater than samplerate/2"
    x, f32, flen, fstep, pre, win = _prepare(signal, samplerate, winlen, winstep, nfft, preemph, winfunc)
    plan = _gpu.mfcc_plan(samplerate=samplerate, frame_len=flen, frame_step=fstep, nfft=nfft, nfilt=nfilt,
                          numcep=min(13, nfilt), preemph=pre, lowfreq=float(lowfreq), highfreq=float(highfreq), window=win)
    off = _gpu.to_device(numpy.array([0, len(x)], dtype=numpy.int64))
    out, _ = plan.fbank_f32(_gpu.to_device(x.astype(numpy.float32)), off)
    n = dspfe.num_frames(len(x), flen, fstep)
    res = out[:n].cpu().numpy().astype(numpy.float64)
    return res[:, :nfilt], res[:, nfilt]


def hz2mel(hz):
    """reference base.py:34-35."""
    return 2595 * numpy.log10(1 + hz / 700.)


def mel2hz(mel):
    """reference base.py:37-38."""
    return 700 * (10 ** (mel / 2595.0) - 1)


def get_filterbanks(nfilt=20, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    """reference base.py:40-58: dense [nfilt, nfft/2+1] triangular filterbank (host table; the kernels build the
    same weights from the same bin edges, see csrc/mfcc_tables.h)."""
    highfreq = highfreq or samplerate / 2
    assert highfreq <= samplerate / 2, "highfreq is greater than samplerate/2"
    melpoints = numpy.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    bin = numpy.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fb = numpy.zeros([nfilt, nfft // 2 + 1])
    for j in range(0, nfilt):
        for i in range(int(bin[j]), int(bin[j + 1])):
            fb[j, i] = (i - bin[j]) / (bin[j + 1] - bin[j])
        for i in range(int(bin[j + 1]), int(bin[j + 2])):
            fb[j, i] = (bin[j + 2] - i) / (bin[j + 2] - bin[j + 1])
    return fb


def lifter(cepstra, L=22):
    """reference base.py:60-68."""
    if L > 0:
        nframes, ncoeff = numpy.shape(cepstra)
        n = numpy.arange(ncoeff)
        return (1 + (L / 2.) * numpy.sin(numpy.pi * n / L)) * cepstra
    return cepstra


def delta(feat, N):
    """reference base.py:70-79, on the device (dspfe_delta_f32).  Keeps the input dtype like numpy.empty_like."""
    if N < 1:
        raise ValueError('N must be an integer >= 1')
    feat = numpy.asarray(feat)
    if feat.ndim != 2:
        raise NotImplementedError("feat must be a 2-D array")
    return dspfe.delta_f32(feat, N).astype(feat.dtype if feat.dtype.kind == 'f' else numpy.float64)
