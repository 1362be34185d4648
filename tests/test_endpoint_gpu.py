"""GPU parity tests of the endpoint kernels (K2a/K2b/K3) through the C ABI: bit-exact integers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def pack(xs):
    off = np.zeros(len(xs) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in xs], out=off[1:])
    return np.concatenate(xs).astype(np.int16), off


def test_golden_batch_bit_exact(golden):
    import torch
    import dspfe
    g = golden("endpoint")
    names = [str(n) for n in g["names"]]
    pcm, off = pack([g[f"{n}/x"] for n in names])
    plan = dspfe.EndpointPlan()
    assert (plan.frame_len, plan.frame_step) == (480, 160)
    lr, asum, zcr, fo = plan.detect_host(pcm, off, want_features=True)
    np.testing.assert_array_equal(lr, g["lr"])
    for i, n in enumerate(names):
        np.testing.assert_array_equal(asum[fo[i]:fo[i + 1]] / 480.0, g[f"{n}/amp"])     # float64 divide == np.mean
        np.testing.assert_array_equal(zcr[fo[i]:fo[i + 1]], g[f"{n}/zcr"])
    dev = torch.device("cuda:0")
    lr_d, asum_d, zcr_d, fo_d = plan.detect(torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev), want_features=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(lr_d.cpu().numpy(), lr)
    np.testing.assert_array_equal(asum_d.cpu().numpy()[: fo[-1]], asum)
    np.testing.assert_array_equal(zcr_d.cpu().numpy()[: fo[-1]], zcr)
    np.testing.assert_array_equal(fo_d.cpu().numpy(), fo)


def test_random_ragged_vs_oracle():
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    lengths = synth.ragged_lengths(192, seed=3)
    lengths[:7] = [1, 479, 480, 481, 640, 8000, 180000]   # the last one has more frames than the decision kernel stages
    pcm, off = synth.synth_batch(lengths, seed0=5000)
    lr, asum, zcr, fo = dspfe.EndpointPlan().detect_host(pcm, off, want_features=True)
    mism = 0
    for u in range(len(lengths)):
        x = pcm[off[u]:off[u + 1]]
        l, r, amp, z = O.basic_endpoint_detection(x, 16000, return_feature=True)
        np.testing.assert_array_equal(asum[fo[u]:fo[u + 1]] / 480.0, np.array(amp))
        np.testing.assert_array_equal(zcr[fo[u]:fo[u + 1]], np.array(z))
        mism += (int(lr[u, 0]), int(lr[u, 1])) != (l, r)
    assert mism == 0, f"{mism} of {len(lengths)} endpoint pairs differ"


@pytest.mark.parametrize("rate,cfg_frame,cfg_step", [(8000, 0.03, 0.01), (16000, 0.025, 0.01), (44100, 0.03, 0.01), (16000, 0.005, 0.01)])
def test_other_rates_and_partial_hops(rate, cfg_frame, cfg_step):
    """frame_len not a multiple of the hop (and shorter than the hop) exercises the partial-block path."""
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    lengths = [int(rate * t) for t in (0.4, 1.0, 2.3, 0.05)]
    pcm, off = synth.synth_batch(lengths, seed0=6000, sr=rate)
    plan = dspfe.EndpointPlan(samplerate=rate, cfg_frame=cfg_frame, cfg_step=cfg_step)
    lr, asum, zcr, fo = plan.detect_host(pcm, off, want_features=True)
    for u in range(len(lengths)):
        x = pcm[off[u]:off[u + 1]]
        l, r, amp, z = O.basic_endpoint_detection(x, rate, return_feature=True, cfg_frame=cfg_frame, cfg_step=cfg_step)
        np.testing.assert_array_equal(asum[fo[u]:fo[u + 1]] / float(plan.frame_len), np.array(amp))
        np.testing.assert_array_equal(zcr[fo[u]:fo[u + 1]], np.array(z))
        assert (int(lr[u, 0]), int(lr[u, 1])) == (l, r)


def test_config4_ragged_endpoint_then_mfcc():
    """BASELINE config 4: endpoint detection + MFCC on a ragged batch, chained on the device through d_trim."""
    import torch
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    from tol import assert_mfcc_close
    dev = torch.device("cuda:0")
    U = 4096
    lengths = synth.ragged_lengths(U, seed=11)
    pcm, off = synth.synth_batch_torch(lengths, seed0=777, device=dev)
    off_d = off.to(dev)
    ep, mf = dspfe.EndpointPlan(), dspfe.MfccPlan(delta_n=3)
    lr = ep.detect(pcm, off_d)
    out, fo = mf.mfcc_delta(pcm, off_d, trim=lr)
    torch.cuda.synchronize()
    lr_h, fo_h, off_h = lr.cpu().numpy(), fo.cpu().numpy(), off.numpy()
    # size-independent properties over the whole batch
    assert np.all(lr_h[:, 0] >= 0) and np.all(lr_h[:, 1] > lr_h[:, 0])
    seg = np.minimum(lr_h[:, 1], lengths) - np.minimum(lr_h[:, 0], lengths)
    np.testing.assert_array_equal(np.diff(fo_h), dspfe.frame_counts(seg, 400, 160))
    assert bool(torch.isfinite(out[: fo_h[-1]]).all())
    # sampled oracle parity: endpoints bit-exact, trimmed MFCC within tolerance
    for u in (0, 1, 2, 1000, 2047, 4095):
        x = pcm[off_h[u]:off_h[u + 1]].cpu().numpy()
        l, r = O.basic_endpoint_detection(x, 16000)
        assert (int(lr_h[u, 0]), int(lr_h[u, 1])) == (l, r)
        assert_mfcc_close(out[fo_h[u]:fo_h[u + 1]].cpu().numpy(), O.mfcc_delta39(x[l:r], 3), what=f"C4 utt {u}")


def test_robust_endpoint_detection_vs_oracle():
    """the autocorrelation-gated rule, bit-exact (integer lag sums)."""
    import torch
    import dspfe
    import features
    from dspfe import synth
    from oracle import ref_features as O
    lengths = synth.ragged_lengths(48, seed=8, lo=8000, hi=48000)
    lengths[:4] = [300, 480, 9000, 170000]
    pcm, off = synth.synth_batch(lengths, seed0=8100)
    plan = dspfe.EndpointPlan()
    lr = plan.detect_robust_host(pcm, off)
    dev = torch.device("cuda:0")
    lr_d = plan.detect_robust(torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(lr_d.cpu().numpy(), lr)
    differs_from_basic = 0
    for u in range(len(lengths)):
        x = pcm[off[u]:off[u + 1]]
        want = O.robust_endpoint_detection(x, 16000)
        assert (int(lr[u, 0]), int(lr[u, 1])) == want, u
        differs_from_basic += want != O.basic_endpoint_detection(x, 16000)
    assert differs_from_basic > 0          # the gate must actually bite on this set
    x = pcm[off[5]:off[6]]
    assert features.robust_endpoint_detection(x, 16000) == O.robust_endpoint_detection(x, 16000)
    # list-typed API: amplitude_rule(use_acr=True, frames=..., rate=...)
    frames = O.to_frames(x, 16000, 0.03, 0.01)
    amp = O.get_amplitude(frames)
    assert features.amplitude_rule(amp, 0.5, frames=frames, use_acr=True, rate=16000) == \
        O.amplitude_rule(amp, 0.5, frames=frames, use_acr=True, rate=16000)


def test_get_amplitude_variants_and_noise():
    import features
    from dspfe import synth
    from oracle import ref_features as O
    x = synth.synth_utterance(77, 9000)
    frames = O.to_frames(x, 16000, 0.03, 0.01)
    for window in ('square', 'hamming'):
        for use_sq in (False, True):
            got = features.get_amplitude(frames, window, use_sq)
            want = O.get_amplitude(frames, window, use_sq)
            assert isinstance(got, list)
            np.testing.assert_allclose(got, want, rtol=1e-12)
    amp = O.get_amplitude(frames)
    sep = O.amplitude_rule(amp)
    assert features.get_noise(amp, sep) == O.get_noise(amp, sep)
    assert features.get_noise(amp, [(0, len(amp))]) == 1e30


def test_warp_parallel_decision_equals_serial_rule():
    """The device decision (rank-sorted silence model, ballot masks, find-first-set walk) against the serial replay of the
    reference rule (dspfe_endpoint_decide_host, the same code the oracle tests pin) on adversarial frame statistics:
    plateaus exactly at the thresholds, all-equal amplitudes, digital silence, one-frame utterances, bursts."""
    import torch
    import dspfe
    from dspfe import synth
    rng = np.random.default_rng(77)
    xs = []
    for u in range(600):
        n = int(rng.integers(1, 60000))
        kind = u % 6
        if kind == 0:
            x = synth.synth_utterance(7000 + u, n)
        elif kind == 1:
            x = (rng.standard_normal(n) * rng.uniform(1, 3000)).astype(np.int16)
        elif kind == 2:
            x = np.full(n, int(rng.integers(-5, 6)), dtype=np.int16)                     # every frame identical (ties everywhere)
        elif kind == 3:
            x = ((rng.integers(0, 2, n) * 2 - 1) * int(rng.integers(1, 20000))).astype(np.int16)   # constant |x|, random signs
        elif kind == 4:
            x = np.zeros(n, dtype=np.int16)
            for _ in range(int(rng.integers(1, 6))):                                      # bursts of equal level
                a = int(rng.integers(0, n)); b = min(n, a + int(rng.integers(1, 9000)))
                x[a:b] = ((rng.integers(0, 2, b - a) * 2 - 1) * 4000).astype(np.int16)
        else:
            x = (np.sin(np.arange(n) * rng.uniform(0.01, 1.5)) * rng.uniform(10, 30000) * (np.arange(n) % 3200 < 1600)).astype(np.int16)
        xs.append(x)
    pcm, off = pack(xs)
    dev = torch.device("cuda:0")
    plan = dspfe.EndpointPlan()
    lr, asum, zcr, fo = plan.detect(torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev), want_features=True)
    torch.cuda.synchronize()
    lr, asum, zcr, fo = lr.cpu().numpy(), asum.cpu().numpy(), zcr.cpu().numpy(), fo.cpu().numpy()
    bad = []
    for u in range(len(xs)):
        want = dspfe.endpoint_decide_host(asum[fo[u]:fo[u + 1]], zcr[fo[u]:fo[u + 1]])
        if tuple(int(v) for v in lr[u]) != tuple(int(v) for v in want):
            bad.append((u, tuple(lr[u]), tuple(want)))
    assert not bad, f"{len(bad)} of {len(xs)} decisions differ, first: {bad[:3]}"


def test_robust_endpoint_detection_other_rates():
    """K3r at 8 kHz (240-sample frames, lags 16..159) and 44.1 kHz (1323-sample frames, lags 88..881: several lag chunks per lane,
    odd frame length): the probe at the predicted lag, the float32 screening and the exact recompute against the oracle."""
    import dspfe
    from dspfe import synth
    from oracle import ref_features as O
    for rate, lo, hi in ((8000, 4000, 30000), (44100, 20000, 120000)):
        lengths = synth.ragged_lengths(10, seed=rate, lo=lo, hi=hi)
        lengths[0] = int(0.03 * rate) - 1
        pcm, off = synth.synth_batch(lengths, seed0=rate + 7, sr=rate)
        lr = dspfe.EndpointPlan(samplerate=rate).detect_robust_host(pcm, off)
        for u in range(len(lengths)):
            want = O.robust_endpoint_detection(pcm[off[u]:off[u + 1]], rate)
            assert (int(lr[u, 0]), int(lr[u, 1])) == want, (rate, u)
