"""GPU parity tests of the pitch kernels (K4a/K4b/K5/K6) through the C ABI.

Contract (BASELINE.json north_star): pitch-peak lags are exact except where the reference statistic lies
within tolerance of the decision threshold; such frames are counted (tests/pitchcheck.py) and every other difference
fails.  pitch_feature floats: 1e-6 relative on utterances whose lags are all exact."""
import numpy as np
import pytest

import pitchcheck

pytestmark = pytest.mark.gpu


def pack(xs):
    off = np.zeros(len(xs) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in xs], out=off[1:])
    return np.concatenate(xs).astype(np.int16), off


def test_golden_pitch_tracks_exact(golden):
    import dspfe
    g = golden("pitch")
    names = ["u0", "u1", "u2", "u3"]
    pcm, off = pack([g[f"{n}/x"] for n in names])
    for method, key, kw in ((0, "pitch_cep", {}), (1, "pitch_sr", {}), (1, "pitch_sr300", dict(frame_len=300))):
        pitch, lag, fo = dspfe.PitchPlan(method=method, **kw).detect_host(pcm, off)
        for i, n in enumerate(names):
            np.testing.assert_array_equal(pitch[fo[i]:fo[i + 1]], g[f"{n}/{key}"], err_msg=f"{n} {key}")


def test_golden_pitch_feature_chain(golden):
    """pitch_model.py:38-41 on the device: endpoints -> pre-emphasis over the whole signal -> slice -> pitch_feature."""
    import torch
    import dspfe
    g = golden("pitch")
    names = ["u0", "u1", "u2", "u3"]
    pcm, off = pack([g[f"{n}/x"] for n in names])
    dev = torch.device("cuda:0")
    pcm_d, off_d = torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev)
    lr = dspfe.EndpointPlan().detect(pcm_d, off_d)
    np.testing.assert_array_equal(lr.cpu().numpy(), np.stack([g[f"{n}/lr"] for n in names]))
    o = dspfe.PitchPlan(method=0, preemph=0.97).detect(pcm_d, off_d, trim=lr, want_feat=True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(o["feat"].cpu().numpy(), g["pitch_feature"], rtol=1e-6, atol=1e-9)
    # host-buffer path gives the same bits
    _, _, _, feat = dspfe.PitchPlan(method=0, preemph=0.97).detect_host(pcm, off, trim=lr.cpu().numpy(), want_feat=True)
    np.testing.assert_array_equal(feat, o["feat"].cpu().numpy())


@pytest.mark.parametrize("method", [0, 1])
def test_ragged_batch_vs_oracle(method):
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    lengths = synth.ragged_lengths(24, seed=21, lo=8000, hi=40000)
    lengths[:6] = [1, 700, 1000, 819, 8001, 110000]   # the last one exceeds the feature kernel's staged frame count
    pcm, off = synth.synth_batch(lengths, seed0=4200)
    pitch, lag, fo = dspfe.PitchPlan(method=method).detect_host(pcm, off)
    tot = np.zeros(3, dtype=np.int64)
    for u in range(len(lengths)):
        tot += pitchcheck.check_lags(method, pcm[off[u]:off[u + 1]], 16000, lag[fo[u]:fo[u + 1]], pitch[fo[u]:fo[u + 1]],
                                     what=f"method {method} utterance {u}")
    pitchcheck.record(f"ragged_batch_vs_oracle[{method}]", *tot)
    assert tot[1] <= 0.02 * tot[0], f"{tot[1]} of {tot[0]} frames are near-ties: the kernel's float32 error is larger than designed"


def test_rows_tap_matches_reference_cepstrum_and_acr():
    import torch
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    x = synth.synth_utterance(99, 24000)
    dev = torch.device("cuda:0")
    xd = torch.from_numpy(x).to(dev)
    od = torch.tensor([0, len(x)], dtype=torch.int64, device=dev)
    fr = O.to_frames(O.downsampling(x, 16000, 10000), 10000, 0.0512, 0.01)
    cl = O.center_clip(fr, False)
    o = dspfe.PitchPlan(method=0, row_len=512).detect(xd, od, want_rows=True)
    torch.cuda.synchronize()
    F = int(o["frame_off"][-1])
    ce = O.pitch_detect_frame(cl, 10000)
    assert F == len(ce)
    assert np.max(np.abs(o["rows"][:F].cpu().numpy() - ce)) <= 1e-5 * np.max(np.abs(ce))
    o = dspfe.PitchPlan(method=1).detect(xd, od, want_rows=True)
    torch.cuda.synchronize()
    sr = O.pitch_detect_frame_sr(cl, 10000)
    assert np.max(np.abs(o["rows"][:F].cpu().numpy() - sr)) <= 2e-5 * np.max(np.abs(sr))


def test_dropin_pitch_module():
    """The reference-facing functions of features.pitch (same names, signatures and return types)."""
    import features
    from dspfe import synth
    from oracle import ref_features as O
    x = synth.synth_utterance(31, 20000)
    l, r = features.basic_endpoint_detection(x, 16000)
    sig = features.preemphasis(x, coeff=0.97)
    feat = features.pitch_feature(sig[l:r], 16000)
    want = O.pitch_feature(O.preemphasis(x, 0.97)[l:r], 16000)
    assert isinstance(feat, tuple) and len(feat) == 5
    np.testing.assert_allclose(feat, want, rtol=1e-6, atol=1e-9)
    p, frames = features.pitch_detect(x, 16000)
    wp, wf = O.pitch_detect(x, 16000)
    assert isinstance(p, list) and p == wp
    np.testing.assert_array_equal(frames, wf)
    p, _ = features.pitch_detect_sr(x, 16000, winlen=0.03, step=0.01)     # as model.py:92 calls it
    assert p == O.pitch_detect_sr(x, 16000, winlen=0.03, step=0.01)[0]
    fr = wf[40]
    np.testing.assert_array_equal(features.center_clip(fr), O.center_clip(fr))
    cc = features.center_clip(fr, False)
    np.testing.assert_allclose(cc, O.center_clip(fr, False), rtol=1e-6, atol=1e-3)
    ce = features.pitch_detect_frame(cc, 10000, 'male')
    wce = O.pitch_detect_frame(O.center_clip(fr, False), 10000)
    assert ce.shape == (512,) and np.max(np.abs(ce - wce)) <= 1e-5 * np.max(np.abs(wce))
    sr = features.pitch_detect_frame_sr(cc[:300], 10000)
    wsr = O.pitch_detect_frame_sr(O.center_clip(fr, False)[:300], 10000)
    assert isinstance(sr, list) and len(sr) == 180
    assert np.max(np.abs(np.array(sr) - np.array(wsr))) <= 2e-5 * np.max(np.abs(wsr))
    rows = O.pitch_detect_frame(O.center_clip(wf[30:50], False), 10000)
    sm = features.smooth(rows.tolist())
    wsm = np.asarray(O.smooth(rows))
    assert isinstance(sm, list) and np.max(np.abs(np.asarray(sm) - wsm)) <= 1e-5 * np.max(np.abs(wsm))
    assert features.peak_score(wsm[5]) == O.peak_score(wsm[5].astype(np.float32))
    sc = [O.peak_score(c) for c in wsm]
    assert features.robust_max_pitch(sc) == O.robust_max_pitch(sc)
    assert features.max_pitch(sc) == O.max_pitch(sc)
    assert features.greedy_max_pitch(sc) == O.greedy_max_pitch(sc)
    seg, idx = features.find_smooth_subsequence(wp, bias=3)
    wseg, widx = O.find_smooth_subsequence(wp, bias=3)
    assert list(seg) == list(wseg) and idx == widx
    np.testing.assert_allclose(features.slope(wseg), O.slope(wseg), rtol=1e-9)
    np.testing.assert_allclose(features.quad_params(wseg), O.quad_params(wseg), rtol=1e-8)
    assert features.peakshift(wseg, wp) == O.peakshift(wseg, wp)
    assert features.sub_endpoint_detect(wf) == O.sub_endpoint_detect(wf)
    assert features.dp_max_pitch(sc) == O.dp_max_pitch(sc)
    # sigproc.window / acr on the device (float64)
    y = features.window(fr, 10000, 50, 1000, 'hamming')
    wy = O.window(fr, 10000, 50, 1000, 'hamming')
    assert y.dtype == np.complex128 and np.max(np.abs(y - wy)) <= 1e-9 * np.max(np.abs(wy))
    y = features.window(fr[:300], 10000, 0, 500)
    assert np.max(np.abs(y - O.window(fr[:300], 10000, 0, 500))) <= 1e-9 * np.max(np.abs(wy))
    for n in (0, 1, 37, 199):
        np.testing.assert_allclose(features.acr(fr, n), O.acr(fr, n), rtol=1e-13)
    import pickle
    assert features.pitch.pickle is pickle    # pitch_model.py star-imports it from here (SURVEY A-12)


def test_config3_full_size_properties():
    """BASELINE config 3: cepstrum + autocorrelation pitch features on 4096 ragged utterances.  Size-independent
    properties over the whole batch plus sampled oracle parity."""
    import torch
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    dev = torch.device("cuda:0")
    U = 4096
    lengths = synth.ragged_lengths(U, seed=33)
    pcm, off = synth.synth_batch_torch(lengths, seed0=31337, device=dev)
    off_d = off.to(dev)
    lr = dspfe.EndpointPlan().detect(pcm, off_d)
    cep = dspfe.PitchPlan(method=0, preemph=0.97)
    o = cep.detect(pcm, off_d, trim=lr, want_feat=True)
    acr = dspfe.PitchPlan(method=1).detect(pcm, off_d)
    torch.cuda.synchronize()
    fo = o["frame_off"].cpu().numpy()
    lr_h, off_h = lr.cpu().numpy(), off.numpy()
    seg = np.minimum(lr_h[:, 1], lengths) - np.minimum(lr_h[:, 0], lengths)
    np.testing.assert_array_equal(np.diff(fo), [cep.num_frames(int(n)) for n in seg])
    lag = o["lag"][: fo[-1]].cpu().numpy()
    pitch = o["pitch"][: fo[-1]].cpu().numpy()
    assert lag.min() >= 20 and lag.max() <= 99
    base = 1.0 / (0.0001 * lag)
    assert np.all((pitch == base) | (pitch == 2 * base))       # robust_max_pitch only ever doubles
    fo2 = acr["frame_off"].cpu().numpy()
    lag2 = acr["lag"][: fo2[-1]].cpu().numpy()
    assert lag2.min() >= 20 and lag2.max() <= 199
    # idempotence / determinism: a second run gives the same bits; a sub-batch equals its slice of the whole
    o2 = cep.detect(pcm, off_d, trim=lr, want_feat=True)
    torch.cuda.synchronize()
    assert torch.equal(o2["pitch"][: fo[-1]], o["pitch"][: fo[-1]])
    assert torch.equal(torch.nan_to_num(o2["feat"]), torch.nan_to_num(o["feat"]))
    a, b = 1000, 1100
    sub = cep.detect(pcm[off_h[a]:off_h[b]].clone(), (off_d[a:b + 1] - off_d[a]).contiguous(), trim=lr[a:b].contiguous(), want_feat=True)
    torch.cuda.synchronize()
    assert torch.equal(sub["pitch"][: fo[b] - fo[a]], o["pitch"][fo[a]:fo[b]])
    # sampled oracle parity
    feat = o["feat"].cpu().numpy()
    tot = np.zeros(3, dtype=np.int64)
    for u in (0, 1, 777, 2048, 4095):
        x = pcm[off_h[u]:off_h[u + 1]].cpu().numpy()
        l, r = int(lr_h[u, 0]), int(lr_h[u, 1])
        sig = O.preemphasis(x, 0.97)[l:r]
        c = pitchcheck.check_lags(0, sig, 16000, lag[fo[u]:fo[u + 1]], pitch[fo[u]:fo[u + 1]], what=f"config 3 utterance {u}")
        tot += c
        if c[1] == 0 and c[0] > 40:
            np.testing.assert_allclose(feat[u], O.pitch_feature(sig, 16000), rtol=1e-6, atol=1e-9)
        tot += pitchcheck.check_lags(1, x, 16000, lag2[fo2[u]:fo2[u + 1]], acr["pitch"][fo2[u]:fo2[u + 1]].cpu().numpy(),
                                     what=f"config 3 autocorrelation utterance {u}")
    pitchcheck.record("config3_full_size", *tot)


@pytest.mark.parametrize("rate", [8000, 44100, 48000])
def test_other_sample_rates(rate):
    """The sample-picking decimator for other source rates (44.1 kHz needs a 441 -> 100 pattern; at 48 kHz the frame's
    source span does not fit the staging buffer and is read in place); 8 kHz is below 10 kHz: every sample is kept."""
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    lengths = [int(rate * t) for t in (0.6, 1.3, 0.2)]
    pcm, off = synth.synth_batch(lengths, seed0=9100, sr=rate)
    for method, fn in ((0, O.pitch_detect), (1, O.pitch_detect_sr)):
        plan = dspfe.PitchPlan(method=method, samplerate=rate)
        plan.reserve(len(lengths), len(pcm))
        pitch, lag, fo = plan.detect_host(pcm, off)
        tot = np.zeros(3, dtype=np.int64)
        for u in range(len(lengths)):
            tot += pitchcheck.check_lags(method, pcm[off[u]:off[u + 1]], rate, lag[fo[u]:fo[u + 1]], pitch[fo[u]:fo[u + 1]],
                                         what=f"rate {rate} method {method} utterance {u}")
        pitchcheck.record(f"other_sample_rates[{rate},{method}]", *tot)


@pytest.mark.parametrize("row_len", [100, 200, 203, 512])
def test_peak_score_adversarial_rows(row_len):
    """peak_score (pitch.py:227-242) through the rows tap (dspfe_track_rows_f32, the same warp routine K4b runs): plateaus,
    constant rows, ramps, isolated spikes, NaN samples and NaN lags, for row lengths with and without a partial last block."""
    import dspfe
    from oracle import ref_features as O
    rng = np.random.default_rng(row_len)
    n = np.arange(row_len, dtype=np.float64)
    rows = [np.ones(row_len), n.copy(), -n, np.cos(n / 7.0), np.floor(np.cos(n / 5.0) * 3) / 3, rng.standard_normal(row_len),
            np.round(rng.standard_normal(row_len) * 2) / 2, np.where(n == 0, 100.0, 0.0), np.where(n == 60, 5.0, np.cos(n / 9.0))]
    r = rng.standard_normal(row_len); r[[3, 41, 77]] = np.nan; rows.append(r)
    r = np.cos(n / 11.0); r[50] = np.nan; rows.append(r)
    r = np.zeros(row_len); r[1] = 1.0; rows.append(r)
    r = np.zeros(row_len); r[row_len - 1] = 1.0; rows.append(r)
    for k in range(40):                                   # smooth random rows of mixed scales: long scans on both sides
        w = int(rng.integers(1, 40))
        rows.append(np.convolve(rng.standard_normal(row_len + w), np.ones(w) / w, mode="valid")[:row_len])
    rows = np.stack(rows).astype(np.float32)
    _, score, lag = dspfe.smooth_rows_f32(rows, mode=0, want_score=True, want_lag=True, do_smooth=False)
    want = np.array([O.peak_score(x) for x in rows])
    np.testing.assert_array_equal(score, want)
    np.testing.assert_array_equal(lag, 20 + np.argmax(want, axis=1))


def test_dp_max_pitch_on_device():
    """dp_max_pitch (pitch.py:208-225) as a device kernel: bit-exact float64 path against the oracle and the
    host replay, including the reference's back-trace start and first-maximum ties."""
    import dspfe
    import features
    from oracle import ref_features as O
    rng = np.random.default_rng(8)
    for rows, cols in ((2, 5), (40, 80), (97, 180), (30, 200)):
        g = rng.standard_normal((rows, cols)) * 20
        g[rows // 2] = np.round(g[rows // 2])            # exact ties
        want = np.array(O.dp_max_pitch(g))
        np.testing.assert_array_equal(dspfe.dp_max_pitch_f64(g), want)
        np.testing.assert_array_equal(dspfe.dp_max_pitch_host(g), want)
        assert features.dp_max_pitch(g.tolist()) == want.tolist()
