"""Shared parity checks (tolerances are BASELINE.json's north_star: identical step counts and diagnostics,
singular values to relative 1e-10, singular vectors to |cos| >= 1 - 1e-8 after sign alignment)."""
import numpy as np

SIGMA_RTOL = 1e-10
COS_TOL = 1e-8


def assert_same_diagnostics(got, ref: dict, fields=("non_zero", "dimensions", "iterations", "transposed",
                                                   "lanczos_steps", "ritz_values_stabilized",
                                                   "significant_values", "singular_values", "random_seed")):
    for f in fields:
        assert getattr(got, f) == ref[f], f"Diagnostics.{f}: got {getattr(got, f)}, oracle {ref[f]}"
    assert got.kappa == ref["kappa"]
    assert list(got.end_interval) == list(ref["end_interval"])


def assert_svd_parity(got, ref, check_vectors=True, cos_tol=COS_TOL, sigma_rtol=SIGMA_RTOL, degenerate_rtol=1e-6):
    """got: svdlibrs_b200.SvdRec, ref: oracle.OracleSvd"""
    assert got.d == ref.d
    assert got.ut.shape == ref.ut.shape and got.vt.shape == ref.vt.shape and got.s.shape == ref.s.shape
    assert_same_diagnostics(got.diagnostics, ref.diagnostics)
    if ref.d == 0:
        return
    rel = np.abs(got.s - ref.s) / np.abs(ref.s)
    assert rel.max() <= sigma_rtol, f"singular values differ: max rel {rel.max():.3e}"
    if check_vectors:
        s = ref.s
        for i in range(ref.d):
            # a vector is only determined up to sign, and only when its singular value is simple
            gap = min([abs(s[i] - s[j]) for j in range(ref.d) if j != i] or [np.inf]) / abs(s[i])
            if gap < degenerate_rtol:
                continue
            for name, g, r in (("u", got.ut[i], ref.ut[i]), ("v", got.vt[i], ref.vt[i])):
                c = abs(np.dot(g, r)) / (np.linalg.norm(g) * np.linalg.norm(r))
                assert c >= 1.0 - cos_tol, f"{name}[{i}]: |cos| = {c!r} (1-cos = {1 - c:.3e}, rel gap {gap:.2e})"


def orthonormality_defect(rows: np.ndarray) -> float:
    g = rows @ rows.T
    return float(np.abs(g - np.eye(rows.shape[0])).max())
