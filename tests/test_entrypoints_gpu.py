"""GPU tests of the reference-facing entry points that the fused-path tests never reach: the drop-in `features` functions
themselves (features.mfcc on the nfft-512 path with int16 and non-integer float input, fbank, powspec / magspec /
logpowspec incl. the truncation case, framesig / to_frames, delta as model.py:76-77 calls it, get_zcr, amplitude_feature,
downsampling), against the live reference's outputs in tests/golden/{helpers,mfcc}.npz.  Every call goes through the C ABI
(dspfe_mfcc_delta[_f32], dspfe_fbank_f32, dspfe_spectrum_f32, dspfe_frames_f64, dspfe_delta_f32, dspfe_row_zcr_f64, ...)."""
import logging

import numpy as np
import pytest

from tol import assert_mfcc_close

pytestmark = pytest.mark.gpu


def rel_close(got, want, tol, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = float(np.max(np.abs(got - want) / (1 + np.abs(want)))) if want.size else 0.0
    assert err <= tol, f"{what}: max |d|/(1+|ref|) = {err:.3e} > {tol:g}"


def test_features_mfcc_int16_and_float(golden):
    """reference base.py:8-16 through features.mfcc: int16 samples take dspfe_mfcc_delta, non-integer float samples
    dspfe_mfcc_delta_f32 (model.py:62-63 hands sklearn-scaled audio to features.mfcc)."""
    import features
    from oracle import ref_features as O
    g = golden("mfcc")
    for name in ("c1_1s", "r_0p5s", "r_1p37s", "short_100", "len_401", "zeros_1000", "fullscale"):
        x = g[f"{name}/x"]
        got = features.mfcc(x)
        assert got.dtype == np.float64
        assert_mfcc_close(got, g[f"{name}/mfcc"], what=f"features.mfcc int16 {name}")
    x = g["r_1p37s/x"]
    xf = x.astype(np.float64) / 1234.5                    # non-integer float input: the float32 sample path
    assert_mfcc_close(features.mfcc(xf), O.mfcc(xf), what="features.mfcc float input")
    xs = O.sk_scale(x.astype(np.float64).reshape(-1, 1), with_mean=False).reshape(-1)    # model.py:62-63
    assert_mfcc_close(features.mfcc(xs), O.mfcc(xs), what="features.mfcc sklearn-scaled input")
    assert_mfcc_close(features.mfcc(x.reshape(1, -1)), g["quirk2d/mfcc"], what="(1,S) pre-emphasis quirk")
    assert_mfcc_close(features.mfcc(x, winfunc=np.hamming), g["hamming/d39_n2"][:, :13], what="hamming window")
    assert_mfcc_close(features.mfcc(x, nfilt=40, numcep=16, ceplifter=0, appendEnergy=False), g["nfilt40_cep20/mfcc"], what="nfilt 40")
    assert_mfcc_close(features.mfcc(x, lowfreq=300, highfreq=3400), g["band/mfcc"], what="band-limited")
    assert_mfcc_close(features.mfcc(x, winlen=0.02, winstep=0.008), g["win20_step8/mfcc"], what="20/8 ms")


def test_features_fbank(golden):
    """reference base.py:18-32 through features.fbank (dspfe_fbank_f32, K1 tap MODE 1)."""
    import features
    g = golden("mfcc")
    feat, energy = features.fbank(g["r_1p37s/x"])
    want_f, want_e = g["fbank/feat"], g["fbank/energy"]
    assert feat.shape == want_f.shape and energy.shape == want_e.shape
    # filterbank energies span many decades: compare relative to each value (float32 arithmetic, 1e-4)
    assert np.max(np.abs(feat - want_f) / np.maximum(np.abs(want_f), 1e-3 * want_f.max())) <= 1e-4
    assert np.max(np.abs(energy - want_e) / np.abs(want_e)) <= 1e-4


def test_features_spectra(golden, caplog):
    """reference sigproc.py:136-175 through features.powspec / magspec / logpowspec (dspfe_spectrum_f32, K1 tap MODE 2)."""
    import features
    from oracle import ref_features as O
    g = golden("helpers")
    fr = g["framesig"]
    scale = float(np.max(g["powspec"]))
    ps = features.powspec(fr, 512)
    assert ps.shape == g["powspec"].shape and np.max(np.abs(ps - g["powspec"])) <= 2e-6 * scale
    ms = features.magspec(fr, 512)
    assert np.max(np.abs(ms - g["magspec"])) <= 2e-6 * float(np.max(g["magspec"]))
    lp = features.logpowspec(fr, 512, norm=0)
    big = g["powspec"] > 1e-6 * scale                      # the log of bins near zero amplifies float32 rounding
    assert np.max(np.abs(lp - g["logpowspec_nonorm"])[big]) <= 1e-3
    lpn = features.logpowspec(fr, 512)
    assert np.max(np.abs(lpn - g["logpowspec"])[big]) <= 1e-3
    # frames longer than NFFT: the reference truncates with a logged warning (sigproc.py:143-146)
    long_fr = np.concatenate([fr, fr[:, :200]], axis=1)   # 600 samples per frame
    with caplog.at_level(logging.WARNING):
        got = features.powspec(long_fr, 512)
    want = O.powspec(long_fr[:, :512], 512)
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= 2e-6 * float(np.max(want))


def test_features_framing(golden):
    """reference sigproc.py:11-19, :66-98 through features.framesig / to_frames (dspfe_frames_f64): bit-exact."""
    import features
    g = golden("helpers")
    x = g["x"]
    np.testing.assert_array_equal(features.framesig(x, 400, 160), g["framesig"])
    np.testing.assert_array_equal(features.framesig(x, 400, 160, np.hamming), g["framesig_ham"])
    np.testing.assert_array_equal(features.to_frames(x, 16000, 0.03, 0.01), g["to_frames_30_10"])
    np.testing.assert_array_equal(features.framesig(x[:100], 400, 160).shape, (1, 400))      # one zero-padded frame
    np.testing.assert_array_equal(features.preemphasis(x), g["preemph_095"])
    np.testing.assert_array_equal(features.preemphasis(x, 0.97), g["preemph_097"])
    np.testing.assert_allclose(features.deframesig(g["framesig"], len(x), 400, 160), g["deframesig"], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("N", [1, 2, 3])
def test_features_delta_standalone(golden, N):
    """reference base.py:70-79 through features.delta (dspfe_delta_f32) on a float64 matrix, as model.py:76-77 calls it:
    delta(mfcc0, 3) and delta(delta(mfcc0, 3), 3)."""
    import features
    from oracle import ref_features as O
    g = golden("mfcc")
    m = g["r_1p37s/mfcc"]
    m0 = m - np.mean(m)                                    # model.py:75
    d1 = features.delta(m0, N)
    assert d1.dtype == np.float64 and d1.shape == m0.shape
    rel_close(d1, O.delta(m0, N), 1e-5, f"delta N={N}")
    d2 = features.delta(d1, N)
    rel_close(d2, O.delta(O.delta(m0, N), N), 2e-5, f"delta-delta N={N}")
    if N in (2, 3):                                        # against the live reference's own chain
        rel_close(features.delta(m, N), g[f"r_1p37s/d39_n{N}"][:, 13:26], 1e-5, f"golden delta N={N}")
    if N == 2:
        np.testing.assert_allclose(features.delta(golden("helpers")["delta_known"] * 0 + np.array([[0.], [1.], [4.], [9.], [16.], [25.]]), 2),
                                   golden("helpers")["delta_known"], rtol=1e-6)
    with pytest.raises(ValueError):
        features.delta(m0, 0)
    # short inputs: fewer rows than the stencil
    for rows in (1, 2, 3):
        rel_close(features.delta(m0[:rows], N), O.delta(m0[:rows], N), 1e-5, f"delta on {rows} rows")


def test_features_zcr_amplitude_downsampling(golden):
    """reference endpoint.py:109-131,182-198 and preprocess.py:21-28 through the drop-in functions."""
    import features
    from oracle import ref_features as O
    ge, gh = golden("endpoint"), golden("helpers")
    for name in ("u0", "u3", "u9", "noise", "click"):
        x = ge[f"{name}/x"]
        fr = features.to_frames(x, 16000, 0.03, 0.01)
        z = features.get_zcr(fr)
        assert isinstance(z, list) and all(isinstance(v, np.int64) for v in z[:3])
        np.testing.assert_array_equal(np.array(z), ge[f"{name}/zcr"])
        a = features.amplitude_feature(x, 16000, 0.03, 0.01)
        assert isinstance(a, list) and isinstance(a[0], np.float64)
        np.testing.assert_array_equal(np.array(a), ge[f"{name}/amp"])
        np.testing.assert_array_equal(np.array(features.get_amplitude(fr)), ge[f"{name}/amp"])
    x = gh["x"]
    np.testing.assert_array_equal(features.downsampling(x, 16000, 10000), gh["downsample_16k_10k"])
    np.testing.assert_array_equal(features.downsampling(x, 44100, 10000), gh["downsample_44k_10k"])
    np.testing.assert_array_equal(features.downsampling(x, 48000, 16000), gh["downsample_48k_16k"])
    # zero-crossing edge cases: zeros never count, the zero-padded tail is part of the last frame
    fr = np.array([[1, -1, 0, 1, -1, -1, 1, 0], [0, 0, 0, 0, 0, 0, 0, 0], [5, -5, 5, -5, 5, -5, 5, -5]], dtype=np.float64)
    np.testing.assert_array_equal(np.array(features.get_zcr(fr)), np.array(O.get_zcr(fr)))


def test_features_endpoint_float_input(golden):
    """reference endpoint.py:34 accepts any real dtype; the drop-in takes float arrays holding int16 values (what
    `wav.read(...).astype(float)` gives) bit-exactly and fractional input through the float64 statistics path."""
    import features
    from oracle import ref_features as O
    g = golden("endpoint")
    x = g["u4/x"]
    assert features.basic_endpoint_detection(x.astype(np.float64), 16000) == tuple(int(v) for v in g["lr"][4])
    xf = x.astype(np.float64) * 0.37
    assert features.basic_endpoint_detection(xf, 16000) == O.basic_endpoint_detection(xf, 16000)
    l, r, amp, zcr = features.basic_endpoint_detection(xf, 16000, return_feature=True)
    wl, wr, wamp, wzcr = O.basic_endpoint_detection(xf, 16000, return_feature=True)
    assert (l, r) == (wl, wr)
    np.testing.assert_allclose(np.array(amp), np.array(wamp), rtol=1e-12)
    np.testing.assert_array_equal(np.array(zcr), np.array(wzcr))


def test_step_longer_than_frame():
    """Gaps between frames (winstep > winlen, cfg.step > cfg.frame): the MFCC plans route such framings to the general
    kernel with a row bound of two frames per utterance (reference sigproc.py:79-87 pads the last frame with zeros)."""
    import dspfe
    import features
    from dspfe import synth
    from oracle import ref_features as O
    x = synth.synth_utterance(5, 16300)
    for kw in (dict(winlen=0.01, winstep=0.025), dict(winlen=0.01, winstep=0.025, nfft=1536), dict(winlen=0.02, winstep=0.0500625, nfft=1024)):
        assert_mfcc_close(features.mfcc(x, **kw), O.mfcc(x, **kw), what=f"gaps {kw}")
    feat, energy = features.fbank(x, winlen=0.01, winstep=0.025)
    wf, we = O.fbank(x, winlen=0.01, winstep=0.025)
    assert feat.shape == wf.shape and np.max(np.abs(feat - wf) / np.maximum(np.abs(wf), 1e-3 * wf.max())) <= 1e-4
    # the endpoint path does take cfg.step > cfg.frame (its frame bound counts two frames per utterance then)
    lengths = [16300, 8000, 16001, 159, 161]
    pcm, off = synth.synth_batch(lengths, seed0=77)
    lr, asum, zcr, fo = dspfe.EndpointPlan(cfg_frame=0.01, cfg_step=0.025).detect_host(pcm, off, want_features=True)
    for u in range(len(lengths)):
        xs = pcm[off[u]:off[u + 1]]
        l, r, amp, z = O.basic_endpoint_detection(xs, 16000, return_feature=True, cfg_frame=0.01, cfg_step=0.025)
        assert (int(lr[u, 0]), int(lr[u, 1])) == (l, r)
        np.testing.assert_array_equal(zcr[fo[u]:fo[u + 1]], np.array(z))
        np.testing.assert_array_equal(asum[fo[u]:fo[u + 1]] / 160.0, np.array(amp))


def test_frames_longer_than_nfft_are_truncated(caplog):
    """reference sigproc.py:143-146: with winlen * samplerate > nfft the frames are counted, padded and windowed at full length
    and the transform takes their first nfft samples (a warning is logged)."""
    import features
    from dspfe import synth
    from oracle import ref_features as O
    x = synth.synth_utterance(88, 20000)
    for kw in (dict(winlen=0.04, winstep=0.01), dict(winlen=0.04, winstep=0.01, winfunc=np.hamming), dict(winlen=0.1, winstep=0.02, nfft=1536)):
        with caplog.at_level(logging.WARNING):
            got = features.mfcc(x, **kw)
        want = O.mfcc(x, **kw)
        assert got.shape == want.shape, (kw, got.shape, want.shape)
        assert_mfcc_close(got, want, what=f"truncated frames {kw}")
    assert any("truncated" in r.message for r in caplog.records)


def test_torch_custom_ops():
    """the batched entry points as torch.library custom ops on CUDA tensors (torch.ops.dspfe.*): same results as
    the plans, fake (shape-only) implementations for tracing, no CPU implementation."""
    import torch
    import dspfe
    import dspfe.torch_ops  # noqa: F401  registers the ops
    from dspfe import synth
    dev = torch.device("cuda:0")
    lengths = [16000, 8000, 30001, 12345]
    pcm, off = synth.synth_batch(lengths, seed0=99)
    pcm_d, off_d = torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev)
    lr = torch.ops.dspfe.endpoint(pcm_d, off_d)
    assert torch.equal(lr, dspfe.EndpointPlan().detect(pcm_d, off_d))
    rows, fo = torch.ops.dspfe.mfcc_delta(pcm_d, off_d, lr)
    want, wfo = dspfe.MfccPlan(delta_n=2).mfcc_delta(pcm_d, off_d, trim=lr)
    n = int(fo[-1])
    assert torch.equal(fo, wfo) and torch.equal(rows[:n], want[:n])
    rows2, fo2 = torch.ops.dspfe.mfcc_delta(pcm_d, off_d, None, 16000, 480, 160, 1536, 3, 0.0, True)      # model.py:74 at 16 kHz
    w2, _ = dspfe.MfccPlan(frame_len=480, frame_step=160, nfft=1536, delta_n=3, preemph=0.0, window=np.hamming(480)).mfcc_delta(pcm_d, off_d)
    assert torch.equal(rows2[: int(fo2[-1])], w2[: int(fo2[-1])])
    hz, lag, pfo = torch.ops.dspfe.pitch(pcm_d, off_d, lr, 1, 16000, 300, 0.0)
    o = dspfe.PitchPlan(method=1, frame_len=300).detect(pcm_d, off_d, trim=lr)
    m = int(pfo[-1])
    assert torch.equal(pfo, o["frame_off"]) and torch.equal(lag[:m], o["lag"][:m]) and torch.equal(hz[:m], o["pitch"][:m])
    # shape-only tracing
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        fp = torch.empty(pcm.shape[0], dtype=torch.int16, device="cuda")
        fo_ = torch.empty(len(lengths) + 1, dtype=torch.int64, device="cuda")
        r, f = torch.ops.dspfe.mfcc_delta(fp, fo_)
        assert r.shape == rows.shape and f.shape == fo.shape
        h, l, q = torch.ops.dspfe.pitch(fp, fo_, None, 1, 16000, 300, 0.0)
        assert h.shape == hz.shape and l.shape == lag.shape
    with pytest.raises(NotImplementedError):
        torch.ops.dspfe.endpoint(torch.from_numpy(pcm), torch.from_numpy(off))
