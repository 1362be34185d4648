"""Pitch-lag parity rule of the contract (BASELINE.json north_star): peak lags are exact, except where the reference
statistic lies within tolerance of the decision threshold -- those frames are counted, and a differing frame that the
float64 oracle does not explain as a near-tie fails the test."""
import json
import os

import numpy as np

REL_TOL = 1e-5          # of max|row|, the float32 error scale of a 512-point transform chain
_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "pitch_mismatch_tests.jsonl")


def check_lags(method, sig, rate, gpu_lag, gpu_pitch=None, winlen=0.0512, step=0.01, what=""):
    """Compares the kernel's lags (20 + argmax, before the octave repair) of one utterance with the float64 oracle.
    Returns (frames, differing, near_tie); asserts that every differing frame is a float64 near-tie and that the Hz track
    is robust_max_pitch of the kernel's own lags, bit for bit."""
    from oracle import ref_features as O
    gpu_lag = np.asarray(gpu_lag)
    if method == 0:
        rows = O.pitch_rows_cep(sig, rate, winlen, step)
        lag = np.array([20 + int(np.argmax(O.peak_score(c))) for c in rows], dtype=np.int64)
        near_fn = O.lag_is_near_tie_cep
    else:
        rows = np.asarray(O.pitch_scores_sr(sig, rate, winlen, step)[0])
        lag = (20 + np.argmax(rows, axis=1)) if len(rows) else np.zeros(0, dtype=np.int64)
        near_fn = O.lag_is_near_tie_sr
    assert len(lag) == len(gpu_lag), f"{what}: {len(gpu_lag)} frames, reference has {len(lag)}"
    bad = np.nonzero(lag != gpu_lag)[0]
    near = [int(i) for i in bad if near_fn(rows[i], gpu_lag[i], REL_TOL)]
    hard = sorted(set(int(i) for i in bad) - set(near))
    assert not hard, (f"{what}: frames {hard[:8]} differ from the float64 reference without a near-tie "
                      f"(kernel lags {gpu_lag[hard[:8]].tolist()}, reference {lag[hard[:8]].tolist()})")
    if gpu_pitch is not None:
        np.testing.assert_array_equal(np.asarray(gpu_pitch), np.asarray(O.robust_pitch_from_lags(gpu_lag)),
                                      err_msg=f"{what}: Hz track is not robust_max_pitch of the kernel's lags")
    return len(lag), len(bad), len(near)


def record(test, frames, differing, near_tie):
    """Appends the counts to gpurun_out/pitch_mismatch_tests.jsonl (kept with the GPU run's artefacts)."""
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps({"test": test, "frames": int(frames), "differing": int(differing), "near_tie": int(near_tie),
                                "hard": int(differing - near_tie)}) + "\n")
    except OSError:
        pass
