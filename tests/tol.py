"""Tolerances of the parity contract (BASELINE.json north_star)."""
import numpy as np

MFCC_TOL = 1e-4   # |gpu - ref| <= MFCC_TOL * (1 + |ref|) against the float64 reference


def mfcc_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.all(np.isfinite(got)), "non-finite output"
    return float(np.max(np.abs(got - ref) / (1.0 + np.abs(ref)))) if got.size else 0.0


def assert_mfcc_close(got, ref, tol=MFCC_TOL, what=""):
    e = mfcc_err(got, ref)
    assert e <= tol, f"{what}: max |d|/(1+|ref|) = {e:.3e} > {tol:.1e}"
    return e
