"""CPU-only checks of the host-side logic inside libdspfe.so (no CUDA call is made):
table construction, frame counting, the endpoint decision rule, and the exported C-ABI symbols."""
import ctypes
import os
import re

import numpy as np
import pytest

import dspfe
from dspfe import synth
from oracle import ref_features as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "dspfe.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(dspfe_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    L = ctypes.CDLL(dspfe.lib_path())
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"declared in include/dspfe.h but not exported: {missing}"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(dspfe.DspfeError) as e:
        dspfe.MfccPlan()
    assert e.value.code == -3
    with pytest.raises(dspfe.DspfeError):
        dspfe.EndpointPlan()


def test_frame_counts_match_reference_rule():
    for flen, step in ((400, 160), (480, 160), (512, 100), (300, 100), (320, 128)):
        for n in list(range(0, 1200)) + [15999, 16000, 16001, 32000, 80000]:
            assert dspfe.num_frames(n, flen, step) == O.num_frames(n, flen, step), (n, flen, step)
    ln = np.array([0, 1, 399, 400, 401, 560, 561, 32000])
    np.testing.assert_array_equal(dspfe.frame_counts(ln, 400, 160), [O.num_frames(int(n), 400, 160) for n in ln])


@pytest.mark.parametrize("cfg", [dict(), dict(nfilt=40), dict(nfilt=20, lowfreq=300, highfreq=3400),
                                 dict(samplerate=8000, nfilt=24), dict(nfilt=13, numcep=13)])
def test_mel_edges_match_reference_filterbank(cfg):
    _, edges = dspfe.mfcc_tables_host(**cfg)
    sr = cfg.get("samplerate", 16000)
    nfilt = cfg.get("nfilt", 26)
    fb = O.get_filterbanks(nfilt, 512, sr, cfg.get("lowfreq", 0), cfg.get("highfreq", None))
    # rebuild the dense filterbank from the edges the kernel uses and compare with the reference matrix
    dense = np.zeros_like(fb)
    for j in range(nfilt):
        lo, mid, hi = edges[j], edges[j + 1], edges[j + 2]
        for i in range(int(lo), int(mid)):
            dense[j, i] = (i - lo) / (mid - lo)
        for i in range(int(mid), int(hi)):
            dense[j, i] = (hi - i) / (hi - mid)
    np.testing.assert_allclose(dense, fb, rtol=0, atol=1e-15)


def test_unsupported_configs_raise():
    # (frame_len > nfft is no longer an error: such frames are truncated for the transform, as the reference does -- sigproc.py:143-146)
    for kw in (dict(nfft=768), dict(nfft=1024), dict(seg_frames=8), dict(nfilt=41), dict(numcep=17), dict(delta_n=0), dict(frame_len=100, frame_step=160)):
        with pytest.raises(dspfe.DspfeError) as e:
            dspfe.mfcc_tables_host(**kw)
        assert e.value.code == -2


def test_endpoint_rule_replay_golden(golden):
    g = golden("endpoint")
    for i, n in enumerate(str(s) for s in g["names"]):
        amp, zcr = g[f"{n}/amp"], g[f"{n}/zcr"]
        asum = np.rint(amp * 480).astype(np.int32)
        np.testing.assert_array_equal(asum / 480.0, amp)          # amp is an exact integer sum over 480
        assert dspfe.endpoint_decide_host(asum, zcr) == tuple(g["lr"][i]), n


def test_endpoint_rule_replay_random():
    """The float64 rule replay in C++ (NumPy summation order included) against the oracle on hostile statistics."""
    rng = np.random.default_rng(5)
    for t in range(400):
        F = int(rng.integers(1, 400))
        kind = t % 4
        if kind == 0:      # plateau of speech on a noise floor
            a = rng.integers(2000, 9000, size=F)
            s, e = sorted(rng.integers(0, F + 1, size=2))
            a[s:e] += rng.integers(0, 400000)
        elif kind == 1:    # pure noise, near-constant: thresholds sit inside the data
            a = rng.integers(10000, 10050, size=F)
        elif kind == 2:    # several bursts
            a = rng.integers(0, 3000, size=F)
            for _ in range(4):
                s = int(rng.integers(0, F)); a[s:s + int(rng.integers(1, 60))] += int(rng.integers(1000, 900000))
        else:              # many exact ties
            a = rng.choice([0, 480, 960, 48000], size=F)
        z = rng.integers(0, 240, size=F)
        amp = [np.float64(v) / 480 for v in a]
        sep = O.amplitude_rule(amp)
        left, right = sep[0][0], sep[-1][1]
        if right - left < 50:
            sep = O.amplitude_rule(amp, 0.125)
        left, right = sep[0][0], sep[-1][1]
        l2, r2 = O.zcr_rule([np.int64(v) for v in z], left, right)
        if r2 - l2 < 50:
            l2, r2 = 0, F
        want = (int(l2 * 0.01 * 16000), int(r2 * 0.01 * 16000))
        assert dspfe.endpoint_decide_host(a, z) == want, (t, F)


def test_sample_index_truncation_quirk():
    """int(k*0.01*16000) != 160*k for some k; the C++ rule must reproduce it."""
    F = 1700
    a = np.full(F, 48000, dtype=np.int32)
    a[:10] = 0; a[-10:] = 0
    z = np.zeros(F, dtype=np.int32)
    l, r = dspfe.endpoint_decide_host(a, z)
    sep = O.amplitude_rule([np.float64(v) / 480 for v in a])
    l2, r2 = O.zcr_rule([np.int64(0)] * F, sep[0][0], sep[-1][1])
    assert (l, r) == (int(l2 * 0.01 * 16000), int(r2 * 0.01 * 16000))
    assert any(int(k * 0.01 * 16000) != 160 * k for k in range(F))


def test_pitch_list_helpers_against_oracle():
    rng = np.random.default_rng(11)
    for t in range(200):
        F = int(rng.integers(1, 300))
        lags = rng.integers(20, 100, size=F)
        if t % 2:   # realistic tracks: a slowly moving lag with octave errors
            base = np.clip(60 + np.cumsum(rng.integers(-2, 3, size=F)), 25, 95)
            lags = np.where(rng.random(F) < 0.15, np.minimum(base * 2, 99), base)
        scores = np.zeros((F, 80), dtype=int)
        scores[np.arange(F), lags - 20] = 5
        want = O.robust_max_pitch(scores)
        got = dspfe.robust_max_pitch_host(lags, repair=True)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(dspfe.robust_max_pitch_host(lags, repair=False), O.max_pitch(scores))
        seg, idx = O.find_smooth_subsequence(want, bias=0)
        gseg, gidx = dspfe.smooth_subsequence_host(want)
        np.testing.assert_array_equal(gseg, seg)
        assert gidx == idx
        amp = rng.random(F) * 1000
        if t % 3 == 0:
            amp = np.round(amp / 100) * 100   # ties
        assert dspfe.sub_endpoint_host(amp) == O.sub_endpoint_detect([[a] for a in amp])
        if len(seg) >= 3:
            np.testing.assert_allclose(dspfe.poly_lead_host(seg, 1), O.slope(seg), rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(dspfe.poly_lead_host(seg, 2), O.quad_params(seg), rtol=1e-8, atol=1e-12)


def test_dp_max_pitch_matches_oracle():
    rng = np.random.default_rng(3)
    for t in range(20):
        g = rng.integers(0, 40, size=(int(rng.integers(2, 60)), 80))
        np.testing.assert_array_equal(dspfe.dp_max_pitch_host(g), O.dp_max_pitch(g))


def test_pitch_frame_counts():
    """decimated length ((S-1)*5-1)//8+2 and the 512/100 framing without a device: the plan
    constructor needs CUDA, so the same formulas are checked through the oracle's index list."""
    for S in (1, 2, 3, 8, 9, 100, 8000, 32000, 80000):
        assert len(O.downsample_indices(S, 16000, 10000)) == ((S - 1) * 5 - 1) // 8 + 2 if S > 1 else 1


def test_torch_custom_ops_register_and_trace_without_a_device():
    """torch.ops.dspfe.* (dspfe/torch_ops.py): registered, shape-only fake implementations agree with the C ABI's bounds, and
    there is no CPU implementation to fall back to."""
    import torch
    import dspfe.torch_ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        pcm = torch.empty(100000, dtype=torch.int16, device="cuda")
        off = torch.empty(5, dtype=torch.int64, device="cuda")
        assert torch.ops.dspfe.endpoint(pcm, off).shape == (4, 2)
        rows, fo = torch.ops.dspfe.mfcc_delta(pcm, off)
        assert rows.shape == (100000 // 160 + 4, 39) and fo.shape == (5,)
        hz, lag, pfo = torch.ops.dspfe.pitch(pcm, off, None, 1, 16000, 300, 0.0)
        assert hz.shape == lag.shape == ((100000 // 8 + 1) * 5 // 100 + 12,) and hz.dtype == torch.float64
    with pytest.raises(NotImplementedError):
        torch.ops.dspfe.endpoint(torch.zeros(10, dtype=torch.int16), torch.zeros(2, dtype=torch.int64))
