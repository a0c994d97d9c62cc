"""Shared helpers for the parity tests."""
import numpy as np


def assert_close(got, want, rtol=1e-4, atol=1e-5, what=""):
    """|got - want| <= atol + rtol*|want| element-wise (north_star: 1e-4 relative / 1e-5 absolute)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    err = np.abs(got - want)
    tol = atol + rtol * np.abs(want)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} elements out of tolerance; worst at {i}: "
                             f"got {got[i]!r} want {want[i]!r} err {err[i]:.3e} tol {tol[i]:.3e}")
    return float(err.max())


def snr_db(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    noise = np.sum((got - want) ** 2)
    return float(10 * np.log10(np.sum(want ** 2) / max(noise, 1e-300)))
