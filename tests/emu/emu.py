"""ctypes binding of the CPU SIMT emulator (tests only; see tests/emu/emu_simt.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libdspfe_emu.so")


class MfccParams(ctypes.Structure):
    _fields_ = [("samplerate", ctypes.c_int32), ("frame_len", ctypes.c_int32), ("frame_step", ctypes.c_int32),
                ("nfft", ctypes.c_int32), ("nfilt", ctypes.c_int32), ("numcep", ctypes.c_int32),
                ("ceplifter", ctypes.c_int32), ("append_energy", ctypes.c_int32), ("delta_n", ctypes.c_int32),
                ("seg_frames", ctypes.c_int32), ("preemph", ctypes.c_double), ("lowfreq", ctypes.c_double),
                ("highfreq", ctypes.c_double), ("window", ctypes.POINTER(ctypes.c_double))]


def build():
    srcs = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(".cpp")]
    csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "dsp-speech-recognition_b200", "csrc")
    deps = srcs + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # DSPFE_EMU_ASAN=1: the kernel bodies under AddressSanitizer (shared-memory carve-ups and workspaces are heap vectors here);
    # run pytest with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 and delete _build/ afterwards
    flags = ["-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if os.environ.get("DSPFE_EMU_ASAN") == "1" else ["-O2"]
    subprocess.check_call(["g++"] + flags + ["-std=c++17", "-DDSPFE_EMU", "-shared", "-fPIC", "-o", LIB] + srcs)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_mfcc_delta.restype = ctypes.c_longlong
        _lib.emu_pitch.restype = ctypes.c_longlong
        _lib.emu_mfcc_long.restype = ctypes.c_longlong
    return _lib


def make_params(frame_len=400, frame_step=160, nfilt=26, numcep=13, ceplifter=22, append_energy=True, delta_n=2,
                seg_frames=0, preemph=0.97, lowfreq=0.0, highfreq=0.0, window=None, samplerate=16000, nfft=512):
    p = MfccParams(samplerate, frame_len, frame_step, nfft, nfilt, numcep, ceplifter, int(append_energy), delta_n,
                   seg_frames, preemph, lowfreq, highfreq, None)
    keep = None
    if window is not None:
        keep = np.ascontiguousarray(window, dtype=np.float64)
        p.window = keep.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    return p, keep


def mfcc_delta(pcm, offsets, trim=None, **kw):
    p, keep = make_params(**kw)
    f32 = np.asarray(pcm).dtype.kind == "f"
    pcm = np.ascontiguousarray(pcm, dtype=np.float32 if f32 else np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    rows = int(len(pcm) // p.frame_step + n)
    out = np.full((rows, 3 * p.numcep), np.nan, dtype=np.float32)
    fo = np.zeros(n + 1, dtype=np.int64)
    err = ctypes.create_string_buffer(256)
    tp = None
    if trim is not None:
        trim = np.ascontiguousarray(trim, dtype=np.int32)
        tp = trim.ctypes.data_as(ctypes.c_void_p)
    r = lib().emu_mfcc_delta(ctypes.byref(p), pcm.ctypes.data_as(ctypes.c_void_p), int(f32), ctypes.c_longlong(len(pcm)),
                             offsets.ctypes.data_as(ctypes.c_void_p), tp, n, out.ctypes.data_as(ctypes.c_void_p),
                             ctypes.c_longlong(rows), fo.ctypes.data_as(ctypes.c_void_p), err, 256)
    if r < 0:
        raise RuntimeError(f"emulator: {err.value.decode()} ({r})")
    return out[:r], fo


def mfcc_long(pcm, offsets, **kw):
    """K1L (nfft = 1536) bodies on the emulator: (out [F, 3*numcep], frame_off)."""
    kw.setdefault("nfft", 1536)
    p, keep = make_params(**kw)
    f32 = np.asarray(pcm).dtype.kind == "f"
    pcm = np.ascontiguousarray(pcm, dtype=np.float32 if f32 else np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    rows = int(len(pcm) // p.frame_step + 2 * n)
    out = np.full((rows, 3 * p.numcep), np.nan, dtype=np.float32)
    fo = np.zeros(n + 1, dtype=np.int64)
    err = ctypes.create_string_buffer(256)
    r = lib().emu_mfcc_long(ctypes.byref(p), pcm.ctypes.data_as(ctypes.c_void_p), int(f32), offsets.ctypes.data_as(ctypes.c_void_p), n,
                            out.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(rows), fo.ctypes.data_as(ctypes.c_void_p), err, 256)
    if r < 0:
        raise RuntimeError(f"emulator: {err.value.decode()} ({r})")
    return out[:r], fo


class PitchParams(ctypes.Structure):
    _fields_ = [("samplerate", ctypes.c_int32), ("dst_rate", ctypes.c_int32), ("frame_len", ctypes.c_int32),
                ("frame_step", ctypes.c_int32), ("method", ctypes.c_int32), ("center_clip", ctypes.c_int32),
                ("row_len", ctypes.c_int32), ("reserved", ctypes.c_int32), ("band_lo", ctypes.c_double),
                ("band_hi", ctypes.c_double), ("preemph", ctypes.c_double)]


def pitch(pcm, offsets, trim=None, method=0, samplerate=16000, dst_rate=10000, frame_len=512, frame_step=100, center_clip=1,
          row_len=0, band_lo=50.0, band_hi=None, preemph=0.0, want_feat=False, want_rows=False):
    """K4/K5/K6 bodies on the emulator.  Returns dict(pitch, lag, frame_off[, feat, rows, smoothed, score])."""
    q = PitchParams(samplerate, dst_rate, frame_len, frame_step, method, center_clip, row_len, 0, band_lo,
                    band_hi if band_hi is not None else (1000.0 if method == 0 else 900.0), preemph)
    f32 = np.asarray(pcm).dtype.kind == "f"
    pcm = np.ascontiguousarray(pcm, dtype=np.float32 if f32 else np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    cap = int(len(pcm) // frame_step + 3 * n + 8)
    rl = (row_len or 200) if method == 0 else 180
    out = dict(pitch=np.full(cap, np.nan), lag=np.zeros(cap, dtype=np.int32), frame_off=np.zeros(n + 1, dtype=np.int64))
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    feat = np.full((n, 5), np.nan) if want_feat else None
    rows = np.full((cap, rl), np.nan, dtype=np.float32) if want_rows else None
    sm = np.full((cap, rl), np.nan, dtype=np.float32) if want_rows else None
    sc = np.zeros((cap, 80), dtype=np.int32) if (want_rows and method == 0) else None
    tp = None if trim is None else np.ascontiguousarray(trim, dtype=np.int32)
    err = ctypes.create_string_buffer(256)
    r = lib().emu_pitch(ctypes.byref(q), vp(pcm), int(f32), vp(offsets), vp(tp), n, vp(out["pitch"]), vp(out["lag"]), vp(feat),
                        vp(rows), vp(sm), vp(sc), vp(out["frame_off"]), ctypes.c_longlong(cap), err, 256)
    if r < 0:
        raise RuntimeError(f"emulator: {err.value.decode()} ({r})")
    out["pitch"], out["lag"] = out["pitch"][:r], out["lag"][:r]
    if want_feat:
        out["feat"] = feat
    if want_rows:
        out["rows"], out["smoothed"] = rows[:r], sm[:r]
        if sc is not None:
            out["score"] = sc[:r]
    return out


def median2(a, b, i16=False):
    """The warp's dual exact median (np.median(frame[frame >= 0]) of two frames at once)."""
    a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape and a.ndim == 1 and len(a) <= 512
    out = np.zeros(2, dtype=np.float32)
    rc = lib().emu_median2(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), len(a), int(i16), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out
