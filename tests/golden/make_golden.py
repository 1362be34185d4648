"""Generate golden fixtures from the LIVE reference (AuCson/DSP-Speech-Recognition `features`).

Run in the CPU container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Inputs are stored with the outputs so the fixtures are
self-contained on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dsp-speech-recognition_b200"))

from oracle import _live_reference as live  # noqa: E402

ref = live.load()
from dspfe import synth  # noqa: E402  (imported after the reference so `features` resolves to the reference)

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}.npz: {os.path.getsize(path) / 1024:.1f} KiB, keys={len(arrs)}")


def mfcc39(x, N, **kw):
    m = ref.mfcc(x, **kw)
    d1 = ref.delta(m, N)
    d2 = ref.delta(d1, N)
    return np.concatenate([m, d1, d2], axis=1)


def gold_mfcc():
    a = {}
    cases = {
        "c1_1s": synth.synth_utterance(101, 16000),          # BASELINE config 1
        "r_0p5s": synth.synth_utterance(102, 8000),
        "r_1p37s": synth.synth_utterance(103, 21931),
        "r_2s": synth.synth_utterance(104, 32000),           # BASELINE config 2 unit
        "short_100": synth.synth_utterance(105, 100),        # one zero-padded frame
        "len_400": synth.synth_utterance(106, 400),
        "len_401": synth.synth_utterance(107, 401),
        "len_561": synth.synth_utterance(108, 561),
        "zeros_1000": np.zeros(1000, dtype=np.int16),        # eps floors (Appendix A-3)
        "fullscale": (np.where(np.arange(3000) % 37 < 18, 32767, -32768)).astype(np.int16),
    }
    for k, x in cases.items():
        a[f"{k}/x"] = x
        with live.quiet():
            a[f"{k}/mfcc"] = ref.mfcc(x)
            a[f"{k}/d39_n2"] = mfcc39(x, 2)
            a[f"{k}/d39_n3"] = mfcc39(x, 3)
    x = cases["r_1p37s"]
    with live.quiet():
        a["hamming/d39_n2"] = mfcc39(x, 2, winfunc=np.hamming)
        a["model_cfg/mfcc2d"] = ref.mfcc(x.reshape(1, -1), 16000, winlen=0.03, winstep=0.01, nfft=512 * 3, winfunc=np.hamming)
        a["quirk2d/mfcc"] = ref.mfcc(x.reshape(1, -1))        # Appendix A-1: equals preemph=0
        a["nopre/mfcc"] = ref.mfcc(x, preemph=0)
        feat, energy = ref.fbank(x)
        a["fbank/feat"], a["fbank/energy"] = feat, energy
        a["nfilt40_cep20/mfcc"] = ref.mfcc(x, nfilt=40, numcep=16, ceplifter=0, appendEnergy=False)
        a["band/mfcc"] = ref.mfcc(x, lowfreq=300, highfreq=3400)
        a["win20_step8/mfcc"] = ref.mfcc(x, winlen=0.02, winstep=0.008)
    save("mfcc", **a)


def gold_helpers():
    a = {}
    x = synth.synth_utterance(201, 4000)
    a["x"] = x
    with live.quiet():
        fr = ref.framesig(x, 400, 160)
        a["framesig"] = fr
        a["framesig_ham"] = ref.framesig(x, 400, 160, np.hamming)
        a["to_frames_30_10"] = ref.to_frames(x, 16000, 0.03, 0.01)
        a["magspec"] = ref.magspec(fr, 512)
        a["powspec"] = ref.powspec(fr, 512)
        a["powspec_trunc256"] = ref.powspec(fr, 256)
        a["logpowspec"] = ref.logpowspec(fr, 512)
        a["logpowspec_nonorm"] = ref.logpowspec(fr, 512, norm=0)
        a["preemph_095"] = ref.preemphasis(x)                 # flat namespace -> preprocess, 0.95
        a["preemph_097"] = ref.preemphasis(x, 0.97)
        a["filterbanks_26_512"] = ref.get_filterbanks(26, 512, 16000)
        a["filterbanks_20_512"] = ref.get_filterbanks()
        a["filterbanks_40_1024_8k"] = ref.get_filterbanks(40, 1024, 8000, 100, 3800)
        a["lifter"] = ref.lifter(np.ones((2, 13)), 22)
        a["delta_known"] = ref.delta(np.array([[0.], [1.], [4.], [9.], [16.], [25.]]), 2)
        a["deframesig"] = ref.deframesig(fr, len(x), 400, 160)
        a["downsample_16k_10k"] = ref.downsampling(x, 16000, 10000)
        a["downsample_44k_10k"] = ref.downsampling(x, 44100, 10000)
        a["downsample_48k_16k"] = ref.downsampling(x, 48000, 16000)
        f10 = ref.to_frames(ref.downsampling(x, 16000, 10000), 10000, 0.0512, 0.01)
        a["pitch_frames"] = f10
        a["center_clip"] = np.stack([ref.center_clip(f, False) for f in f10])
        a["center_clip_bin"] = np.stack([ref.center_clip(f, True) for f in f10])
        a["window_50_1000"] = np.stack([ref.window(f, 10000, 50, 1000, "hamming") for f in f10[:4]])
        a["window_300"] = ref.window(f10[0][:300], 10000, 50, 900, "hamming")
        cep = np.stack([ref.pitch_detect_frame(ref.center_clip(f, False), 10000, "male") for f in f10])
        a["cepstrum"] = cep
        a["smooth_cep"] = np.array(ref.smooth(cep))
        a["peak_score"] = np.array([ref.peak_score(c) for c in ref.smooth(cep)])
        sr = np.stack([ref.pitch_detect_frame_sr(ref.center_clip(f, False), 10000) for f in f10])
        a["acr_scores"] = sr
        a["smooth_9x3"] = np.array(ref.smooth(np.arange(27.0).reshape(9, 3) ** 2))
        a["robust_max_pitch"] = np.array(ref.robust_max_pitch(a["peak_score"].tolist()))
        a["acr_5"] = np.array([ref.acr(f10[1], n) for n in (0, 1, 20, 199)])
    save("helpers", **a)


def gold_endpoint():
    a = {}
    lengths = [8000, 12345, 16000, 21931, 32000, 32000, 40000, 47777, 64000, 80000, 9000, 25000]
    lr, nfr = [], []
    for i, n in enumerate(lengths):
        x = synth.synth_utterance(300 + i, n)
        with live.quiet():
            l, r, amp, zcr = ref.basic_endpoint_detection(x, 16000, return_feature=True)
        a[f"u{i}/x"] = x
        a[f"u{i}/amp"] = np.array(amp, dtype=np.float64)
        a[f"u{i}/zcr"] = np.array(zcr, dtype=np.int64)
        lr.append((l, r))
    # hostile cases: pure noise (fallback to whole signal), click, digital silence head
    rng = np.random.default_rng(7)
    extra = {
        "noise": (rng.standard_normal(20000) * 200).astype(np.int16),
        "click": np.concatenate([np.zeros(5000), [30000, -30000] * 40, np.zeros(9000)]).astype(np.int16) + (rng.standard_normal(14080) * 5).astype(np.int16),
        "late": np.concatenate([(rng.standard_normal(30000) * 30).astype(np.int16), synth.synth_utterance(399, 20000)]),
    }
    for k, x in extra.items():
        with live.quiet():
            l, r, amp, zcr = ref.basic_endpoint_detection(x, 16000, return_feature=True)
        a[f"{k}/x"], a[f"{k}/amp"], a[f"{k}/zcr"] = x, np.array(amp), np.array(zcr, dtype=np.int64)
        lr.append((l, r))
    a["lr"] = np.array(lr, dtype=np.int64)
    a["names"] = np.array([f"u{i}" for i in range(len(lengths))] + list(extra))
    save("endpoint", **a)


def gold_pitch():
    a = {}
    feats = []
    for i, n in enumerate([16000, 24000, 32000, 40000]):
        x = synth.synth_utterance(500 + i, n)
        with live.quiet():
            pc, fr = ref.pitch_detect(x, 16000)
            ps, _ = ref.pitch_detect_sr(x, 16000)
            ps300, _ = ref.pitch_detect_sr(x, 16000, winlen=0.03, step=0.01)   # as model.py:92 calls it
            l, r = ref.basic_endpoint_detection(x, 16000)
            y = ref.preemphasis(x, coeff=0.97)                                  # pitch_model.py:39-41
            feats.append(ref.pitch_feature(y[l:r], 16000))
        a[f"u{i}/x"] = x
        a[f"u{i}/pitch_cep"] = np.array(pc)
        a[f"u{i}/pitch_sr"] = np.array(ps)
        a[f"u{i}/pitch_sr300"] = np.array(ps300)
        a[f"u{i}/lr"] = np.array([l, r])
    a["pitch_feature"] = np.array(feats, dtype=np.float64)
    save("pitch", **a)


def gold_model():
    """SURVEY row f-1 pinned on the live caller: model.py's own _ModelBase.endpoint_detect (:52-64), feature_extract_mfcc
    (:66-88) and pad_batch (:35-50) run on synthetic utterances; the batch is assembled as get_batch_full (:114-135) does
    (its last two lines only move the arrays into a float32 torch tensor [T, B, 39])."""
    import os as _os
    cwd = _os.getcwd()
    _os.chdir(_os.environ.get("TMPDIR", "/tmp"))          # config.py writes ./log at import
    try:
        with live.quiet():
            import model as ref_model                      # the reference's model.py (torch imports included)
    finally:
        _os.chdir(cwd)
    mb = object.__new__(ref_model._ModelBase)              # __init__ wants the data directory; the methods do not
    a = {}
    lengths = [16000, 52000, 9000, 33333, 70000, 24000, 120000]      # the last one keeps more than 200 frames: truncation
    sounds, feats, len0 = [], [], []
    for i, n in enumerate(lengths):
        x = synth.synth_utterance(640 + i, n)
        a[f"u{i}/x"] = x
        with live.quiet():
            sound = mb.endpoint_detect(x, 16000)
            (m0, m1, m2), l = mb.feature_extract_mfcc(sound, 16000)
        feats.append((m0, m1, m2)); len0.append(l)
    with live.quiet():
        cols = [mb.pad_batch(list(c)).transpose(1, 0, 2) for c in zip(*feats)]      # [T, B, 13] each
    a["inp"] = np.concatenate(cols, axis=2).astype(np.float64)
    a["len0"] = np.array(len0)
    a["n"] = np.array(len(lengths))
    save("model", **a)


if __name__ == "__main__":
    gold_model()
    gold_mfcc()
    gold_helpers()
    gold_endpoint()
    gold_pitch()
