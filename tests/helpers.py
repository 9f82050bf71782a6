"""Shared comparison helpers for the parity tests."""
import numpy as np


def check_packed(g: dict, name: str, arr, rtol: float, what: str = ""):
    """Compare `arr` with the fixture entry `name` written by oracle/make_golden.py:pack —
    either the full tensor or (strided subsample, L2 norm, sum). Tolerance is norm-relative:
    max|a-b| <= rtol * max|b| (elementwise relative error is meaningless near zero)."""
    arr = np.asarray(arr, dtype=np.float64)
    if name in g:
        ref = g[name]
        assert arr.shape == ref.shape, f"{what}{name}: shape {arr.shape} vs {ref.shape}"
        scale = max(np.abs(ref).max(), 1e-300)
        err = np.abs(arr - ref).max() / scale
        assert err <= rtol, f"{what}{name}: norm-rel err {err:.3e} > {rtol:.1e}"
        return err
    ref = g[name + "__sub"]
    stride = int(g[name + "__stride"])
    assert tuple(g[name + "__shape"]) == arr.shape, f"{what}{name}: shape {arr.shape} vs {tuple(g[name + '__shape'])}"
    sub = arr.ravel()[::stride]
    scale = max(np.abs(ref).max(), 1e-300)
    err = np.abs(sub - ref).max() / scale
    assert err <= rtol, f"{what}{name}: subsample norm-rel err {err:.3e} > {rtol:.1e}"
    nref = float(g[name + "__norm"])
    nerr = abs(np.sqrt(np.sum(arr * arr)) - nref) / max(nref, 1e-300)
    assert nerr <= rtol, f"{what}{name}: L2 norm rel err {nerr:.3e} > {rtol:.1e}"
    return max(err, nerr)


def nrel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"shape {a.shape} vs {b.shape}"
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
