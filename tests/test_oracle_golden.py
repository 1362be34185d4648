"""The oracle (oracle/ref_features.py) against the golden outputs of the LIVE reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import ref_features as O

TIGHT = dict(rtol=1e-9, atol=1e-9)


def test_mfcc_cases(golden):
    g = golden("mfcc")
    names = sorted({k.split("/")[0] for k in g.files if k.endswith("/x")})
    assert "c1_1s" in names and "zeros_1000" in names
    for n in names:
        x = g[f"{n}/x"]
        np.testing.assert_allclose(O.mfcc(x), g[f"{n}/mfcc"], **TIGHT)
        np.testing.assert_allclose(O.mfcc_delta39(x, 2), g[f"{n}/d39_n2"], **TIGHT)
        np.testing.assert_allclose(O.mfcc_delta39(x, 3), g[f"{n}/d39_n3"], **TIGHT)
    assert g["c1_1s/mfcc"].shape == (99, 13)          # BASELINE config 1
    assert g["r_2s/d39_n2"].shape == (199, 39)        # BASELINE config 2 unit
    z = g["zeros_1000/mfcc"]
    assert abs(z[0, 0] - np.log(np.finfo(float).eps)) < 1e-12 and np.all(np.abs(z[:, 1:]) < 1e-9)


def test_mfcc_variants(golden):
    g = golden("mfcc")
    x = g["r_1p37s/x"]
    np.testing.assert_allclose(O.mfcc_delta39(x, 2, winfunc=np.hamming), g["hamming/d39_n2"], **TIGHT)
    np.testing.assert_allclose(O.mfcc(x.reshape(1, -1), 16000, winlen=0.03, winstep=0.01, nfft=1536, winfunc=np.hamming),
                               g["model_cfg/mfcc2d"], **TIGHT)
    np.testing.assert_allclose(O.mfcc(x.reshape(1, -1)), g["quirk2d/mfcc"], **TIGHT)
    np.testing.assert_allclose(g["quirk2d/mfcc"], g["nopre/mfcc"], rtol=0, atol=0)   # Appendix A-1
    feat, energy = O.fbank(x)
    np.testing.assert_allclose(feat, g["fbank/feat"], **TIGHT)
    np.testing.assert_allclose(energy, g["fbank/energy"], **TIGHT)
    np.testing.assert_allclose(O.mfcc(x, nfilt=40, numcep=16, ceplifter=0, appendEnergy=False), g["nfilt40_cep20/mfcc"], **TIGHT)
    np.testing.assert_allclose(O.mfcc(x, lowfreq=300, highfreq=3400), g["band/mfcc"], **TIGHT)
    np.testing.assert_allclose(O.mfcc(x, winlen=0.02, winstep=0.008), g["win20_step8/mfcc"], **TIGHT)


def test_helpers(golden):
    g = golden("helpers")
    x = g["x"]
    fr = O.framesig(x, 400, 160)
    np.testing.assert_array_equal(fr, g["framesig"])
    np.testing.assert_allclose(O.framesig(x, 400, 160, np.hamming), g["framesig_ham"], **TIGHT)
    np.testing.assert_array_equal(O.to_frames(x, 16000, 0.03, 0.01), g["to_frames_30_10"])
    np.testing.assert_allclose(O.magspec(fr, 512), g["magspec"], **TIGHT)
    np.testing.assert_allclose(O.powspec(fr, 512), g["powspec"], **TIGHT)
    np.testing.assert_allclose(O.powspec(fr, 256), g["powspec_trunc256"], **TIGHT)
    np.testing.assert_allclose(O.logpowspec(fr, 512), g["logpowspec"], **TIGHT)
    np.testing.assert_allclose(O.logpowspec(fr, 512, norm=0), g["logpowspec_nonorm"], **TIGHT)
    np.testing.assert_allclose(O.preemphasis(x), g["preemph_095"], **TIGHT)
    np.testing.assert_allclose(O.preemphasis(x, 0.97), g["preemph_097"], **TIGHT)
    np.testing.assert_allclose(O.get_filterbanks(26, 512, 16000), g["filterbanks_26_512"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(O.get_filterbanks(), g["filterbanks_20_512"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(O.get_filterbanks(40, 1024, 8000, 100, 3800), g["filterbanks_40_1024_8k"], rtol=0, atol=1e-15)
    assert int((g["filterbanks_26_512"] != 0).sum()) == 459
    np.testing.assert_allclose(O.lifter(np.ones((2, 13)), 22), g["lifter"], **TIGHT)
    np.testing.assert_allclose(O.delta(np.array([[0.], [1.], [4.], [9.], [16.], [25.]]), 2), g["delta_known"], **TIGHT)
    np.testing.assert_allclose(g["delta_known"][:, 0], [0.9, 2.2, 4.0, 6.0, 5.8, 4.1], **TIGHT)  # Appendix A-5
    np.testing.assert_allclose(O.deframesig(fr, len(x), 400, 160), g["deframesig"], **TIGHT)
    with pytest.raises(ValueError):
        O.delta(np.zeros((3, 2)), 0)


def test_downsampling(golden):
    g = golden("helpers")
    x = g["x"]
    np.testing.assert_array_equal(O.downsampling(x, 16000, 10000), g["downsample_16k_10k"])
    np.testing.assert_array_equal(O.downsampling(x, 44100, 10000), g["downsample_44k_10k"])
    np.testing.assert_array_equal(O.downsampling(x, 48000, 16000), g["downsample_48k_16k"])
    # closed forms quoted in SURVEY §8 a14
    idx = O.downsample_indices(32000, 16000, 10000)
    assert len(idx) == ((32000 - 1) * 5 - 1) // 8 + 2 == 20001
    k = np.arange(1, len(idx))
    np.testing.assert_array_equal(idx[1:], (8 * (k - 1)) // 5 + 1)


def test_pitch_blocks(golden):
    g = golden("helpers")
    f10 = g["pitch_frames"]
    np.testing.assert_array_equal(O.to_frames(O.downsampling(g["x"], 16000, 10000), 10000, 0.0512, 0.01), f10)
    cc = O.center_clip(f10, False)
    np.testing.assert_array_equal(cc, g["center_clip"])
    np.testing.assert_array_equal(O.center_clip(f10, True), g["center_clip_bin"])
    np.testing.assert_allclose(O.window(f10[:4], 10000, 50, 1000, "hamming"), g["window_50_1000"], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(O.window(f10[0][:300], 10000, 50, 900, "hamming"), g["window_300"], rtol=1e-9, atol=1e-7)
    cep = O.pitch_detect_frame(cc, 10000)
    np.testing.assert_allclose(cep, g["cepstrum"], rtol=1e-7, atol=1e-9)
    sm = np.array(O.smooth(g["cepstrum"]))
    np.testing.assert_allclose(sm, g["smooth_cep"], **TIGHT)
    np.testing.assert_array_equal(np.array([O.peak_score(c) for c in g["smooth_cep"]]), g["peak_score"])
    np.testing.assert_allclose(O.pitch_detect_frame_sr(cc, 10000), g["acr_scores"], rtol=1e-7, atol=1e-6)
    np.testing.assert_allclose(np.array(O.smooth(np.arange(27.0).reshape(9, 3) ** 2)), g["smooth_9x3"], **TIGHT)
    np.testing.assert_allclose(O.robust_max_pitch(g["peak_score"].tolist()), g["robust_max_pitch"], **TIGHT)
    np.testing.assert_allclose([O.acr(f10[1], n) for n in (0, 1, 20, 199)], g["acr_5"], **TIGHT)


def test_endpoint(golden):
    g = golden("endpoint")
    names = [str(n) for n in g["names"]]
    assert len(names) == 15
    for i, n in enumerate(names):
        x = g[f"{n}/x"]
        l, r, amp, zcr = O.basic_endpoint_detection(x, 16000, return_feature=True)
        assert (l, r) == tuple(g["lr"][i]), n
        assert isinstance(amp, list) and isinstance(zcr, list)
        np.testing.assert_array_equal(np.array(amp), g[f"{n}/amp"])      # bit-exact (int16 input)
        np.testing.assert_array_equal(np.array(zcr), g[f"{n}/zcr"])


def test_pitch_end_to_end(golden):
    g = golden("pitch")
    for i in range(4):
        x = g[f"u{i}/x"]
        pc, _ = O.pitch_detect(x, 16000)
        np.testing.assert_allclose(pc, g[f"u{i}/pitch_cep"], **TIGHT)
        ps, _ = O.pitch_detect_sr(x, 16000)
        np.testing.assert_allclose(ps, g[f"u{i}/pitch_sr"], **TIGHT)
        ps3, _ = O.pitch_detect_sr(x, 16000, winlen=0.03, step=0.01)
        np.testing.assert_allclose(ps3, g[f"u{i}/pitch_sr300"], **TIGHT)
        l, r = O.basic_endpoint_detection(x, 16000)
        assert (l, r) == tuple(g[f"u{i}/lr"])
        y = O.preemphasis(x, 0.97)
        np.testing.assert_allclose(O.pitch_feature(y[l:r], 16000), g["pitch_feature"][i], rtol=1e-7, atol=1e-9)


def test_model_batch_pinned_on_live_model_py(golden):
    """oracle.model_batch (restated model.py:35-88,114-135 glue) against the live model.py's output."""
    g = golden("model")
    utts = [g[f"u{i}/x"] for i in range(int(g["n"]))]
    inp, len0 = O.model_batch(utts, N=3, winlen=0.03, winstep=0.01, nfft=1536, preemph=0, winfunc=np.hamming)
    np.testing.assert_array_equal(len0, g["len0"])
    np.testing.assert_allclose(inp, g["inp"], rtol=0, atol=1e-11)
