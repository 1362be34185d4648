"""K4/K5/K6 kernel bodies (csrc/pitch_kernel.cuh) executed on the CPU SIMT emulator (tests/emu) against the golden
vectors of the live reference and against the oracle.  Checks decimation indices, framing, the exact median, the FFT
chains, the smoothing recurrence, peak scoring and the feature tail without a GPU; the GPU parity tests proper are
tests/test_pitch_gpu.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import emu  # noqa: E402
from dspfe import synth  # noqa: E402
from oracle import ref_features as O  # noqa: E402


def test_golden_cepstrum_pitch_and_feature(golden):
    g = golden("pitch")
    for i, name in enumerate(("u0", "u1")):
        x = g[f"{name}/x"]
        r = emu.pitch(x, [0, len(x)], method=0)
        np.testing.assert_array_equal(r["pitch"], g[f"{name}/pitch_cep"])
        # the pitch_model.py call (:38-41): endpoints, pre-emphasis over the whole signal, slice, pitch_feature
        r = emu.pitch(x, [0, len(x)], trim=g[f"{name}/lr"].reshape(1, 2), preemph=0.97, method=0, want_feat=True)
        np.testing.assert_allclose(r["feat"][0], g["pitch_feature"][i], rtol=1e-6, atol=1e-9)


def test_golden_autocorrelation_pitch(golden):
    g = golden("pitch")
    x = g["u0/x"]
    np.testing.assert_array_equal(emu.pitch(x, [0, len(x)], method=1)["pitch"], g["u0/pitch_sr"])
    np.testing.assert_array_equal(emu.pitch(x, [0, len(x)], method=1, frame_len=300)["pitch"], g["u0/pitch_sr300"])


def test_rows_smoothing_and_scores_against_oracle():
    x = synth.synth_utterance(321, 9000)
    r = emu.pitch(x, [0, len(x)], method=0, want_rows=True, row_len=512)
    xs = O.downsampling(x, 16000, 10000)
    fr = O.to_frames(xs, 10000, 0.0512, 0.01)
    ce = O.pitch_detect_frame(O.center_clip(fr, False), 10000)
    assert r["rows"].shape == ce.shape
    assert np.max(np.abs(r["rows"] - ce)) <= 1e-5 * np.max(np.abs(ce))
    sm = np.asarray(O.smooth(ce))
    assert np.max(np.abs(r["smoothed"] - sm)) <= 1e-5 * np.max(np.abs(sm))
    sc = np.asarray([O.peak_score(c) for c in sm])
    assert np.mean(sc != r["score"]) < 0.002   # integer scores; a float32 near-tie may move one
    np.testing.assert_array_equal(r["pitch"], O.robust_max_pitch(sc))


def test_ragged_batch_trim_preemphasis():
    """The pitch_model.py call: preemphasis(sig, 0.97) over the whole utterance, then sig[l:r] -> pitch_feature."""
    lengths = [12000, 5000, 300, 20011]
    pcm, off = synth.synth_batch(lengths, seed0=77)
    trim = np.array([[1000, 11000], [0, 99999], [10, 200], [3333, 18000]], dtype=np.int32)
    r = emu.pitch(pcm, off, trim=trim, method=0, preemph=0.97, want_feat=True)
    for u in range(len(lengths)):
        x = pcm[off[u]:off[u + 1]]
        sig = O.preemphasis(x, 0.97)[trim[u, 0]:trim[u, 1]]
        want, frames = O.pitch_detect(sig, 16000)
        got = r["pitch"][r["frame_off"][u]:r["frame_off"][u + 1]]
        assert len(got) == len(want)
        assert np.mean(got != np.asarray(want)) <= 0.02, u
        if len(want) >= 40 and np.array_equal(got, want):
            np.testing.assert_allclose(r["feat"][u], O.pitch_feature(sig, 16000), rtol=1e-6, atol=1e-9)


def test_edge_cases_zero_short_and_single_frame():
    # digital silence: log 0 -> NaN cepstrum -> every peak score 0 -> lag 20 -> 500 Hz (SURVEY A-11)
    z = np.zeros(3000, dtype=np.int16)
    r = emu.pitch(z, [0, len(z)], method=0)
    want, _ = O.pitch_detect(z, 16000)
    np.testing.assert_array_equal(r["pitch"], want)
    # one zero-padded frame (F = 1: smooth divides an empty window -> NaN row), two frames, 16-byte-unaligned starts
    pcm, off = synth.synth_batch([700, 1000, 1], seed0=5)
    r = emu.pitch(pcm, off, method=0)
    for u in range(3):
        want, _ = O.pitch_detect(pcm[off[u]:off[u + 1]], 16000)
        np.testing.assert_array_equal(r["pitch"][r["frame_off"][u]:r["frame_off"][u + 1]], want)
    r = emu.pitch(pcm, off, method=1)
    for u in range(3):
        want, _ = O.pitch_detect_sr(pcm[off[u]:off[u + 1]], 16000)
        np.testing.assert_array_equal(r["pitch"][r["frame_off"][u]:r["frame_off"][u + 1]], want)


@pytest.mark.parametrize("rate", [8000, 16000, 22050, 44100, 48000])
def test_decimator_pattern_matches_reference(rate):
    n = 2000
    x = np.arange(n, dtype=np.float32)   # sample value = its own index
    r = emu.pitch(x, [0, n], method=1, samplerate=rate, frame_len=300, center_clip=0, want_rows=True)
    idx = O.downsample_indices(n, rate, 10000)
    fr = O.framesig(idx.astype(np.float64), 300, 100)
    assert len(r["pitch"]) == len(fr)
    want = O.pitch_detect_frame_sr(fr, 10000)          # the gathered samples are the kept indices themselves
    assert np.max(np.abs(r["rows"] - want)) <= 2e-5 * np.max(np.abs(want))


@pytest.mark.parametrize("frame_len", [300, 313, 256, 200])
def test_short_frame_autocorrelation_quads(frame_len):
    """pitch_acr_quad (four frames per warp, the split FIR and the packed autocorrelation): the rows of every frame count
    modulo 4 against pitch_detect_frame_sr of the oracle, for frame lengths on both sides of the tap split (T = 512 - L)."""
    lengths = [9000, 4961, 5121, 5281, 700]          # 54, 29, 30, 31 frames of 300 samples (+ a one-frame utterance)
    pcm, off = synth.synth_batch(lengths, seed0=31)
    r = emu.pitch(pcm, off, method=1, frame_len=frame_len, want_rows=True)
    for u in range(len(lengths)):
        x = pcm[off[u]:off[u + 1]]
        fr = O.to_frames(O.downsampling(x, 16000, 10000), 10000, frame_len / 10000.0, 0.01)
        want = np.asarray(O.pitch_detect_frame_sr(O.center_clip(fr, False), 10000))
        got = r["rows"][r["frame_off"][u]:r["frame_off"][u + 1]]
        assert got.shape == want.shape, (u, got.shape, want.shape)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        # float32 transforms: absolute error ~1e-7 of sum v^2, and the last lags divide it by L - n -> 1
        assert np.max(np.abs(got[fin] - want[fin])) <= 5e-5 * np.max(np.abs(want[fin])), u
        p_want, _ = O.pitch_detect_sr(x, 16000, winlen=frame_len / 10000.0, step=0.01)
        assert np.mean(r["pitch"][r["frame_off"][u]:r["frame_off"][u + 1]] != np.asarray(p_want)) <= 0.02


def test_preemphasis_zeros_keep_the_median_sample_set():
    """x[n] = 97 j after x[n-1] = 100 j gives x[n] - 0.97 x[n-1] == 0.0 exactly in the reference's float64 (a member of
    center_clip's `frame >= 0` set, pitch.py:146); a float32 evaluation leaves +-1e-13 there, flips the parity of the set and
    moves the median by a whole order statistic (found on the 8192-utterance bench batch: cepstrum rows off by 1e-3).  The kernel
    evaluates the pre-emphasis in float64 with NumPy's two roundings."""
    rng = np.random.default_rng(3)
    x = synth.synth_utterance(77, 16000).astype(np.int64)
    for k in rng.integers(200, 15800, size=400):                # plant exact zeros of the pre-emphasised signal
        j = int(rng.integers(-80, 80))
        x[k - 1], x[k] = 100 * j, 97 * j
    x = x.astype(np.int16)
    pre = O.preemphasis(x, 0.97)
    assert np.sum(pre == 0.0) >= 300
    r = emu.pitch(x, [0, len(x)], method=0, preemph=0.97, want_rows=True, row_len=512)
    fr = O.to_frames(O.downsampling(pre, 16000, 10000), 10000, 0.0512, 0.01)
    ce = O.pitch_detect_frame(O.center_clip(fr, False), 10000)
    assert np.max(np.abs(r["rows"] - ce)) <= 1e-5 * np.max(np.abs(ce))
    np.testing.assert_array_equal(r["pitch"], O.pitch_detect(pre, 16000)[0])


@pytest.mark.parametrize("i16", [False, True])
def test_dual_median_adversarial(i16):
    """The warp's exact dual median (bit-sliced radix select) against np.median(frame[frame >= 0]): ties, duplicates around the
    middle, zeros and negative zeros, even / odd counts, no non-negative sample, short frames."""
    rng = np.random.default_rng(11 + i16)
    def ref(v):
        nn = v[v >= 0]
        return np.float32(np.median(nn.astype(np.float64))) if len(nn) else np.float32(np.nan)
    cases = []
    for L in (512, 511, 300, 33, 1):
        for _ in range(6):
            if i16:
                a = rng.integers(-32768, 32768, size=L).astype(np.float32)
                b = (rng.integers(-50, 50, size=L) * rng.integers(0, 3, size=L)).astype(np.float32)      # many duplicates and zeros
            else:
                a = (rng.standard_normal(L) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
                b = np.round(rng.standard_normal(L) * 3).astype(np.float32) / 4
                b[rng.integers(0, L, size=max(L // 8, 1))] = -0.0
            cases.append((a, b))
        cases.append((-np.ones(L, dtype=np.float32), np.zeros(L, dtype=np.float32)))                      # empty set / all zeros
        cases.append((np.full(L, 7, dtype=np.float32), np.arange(L, dtype=np.float32) - L // 2))
    for a, b in cases:
        got = emu.median2(a, b, i16=i16)
        for g, w in zip(got, (ref(a), ref(b))):
            assert (np.isnan(g) and np.isnan(w)) or g == w, (len(a), g, w)
