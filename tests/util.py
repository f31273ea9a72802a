"""Comparison helpers shared by the parity tests."""
import numpy as np


def f16(a):
    return np.asarray(a).view(np.float16).astype(np.float32)


def ulp16_diff(a, b):
    """Distance in fp16 representable steps between two uint16 bit-pattern arrays (NaN == NaN -> 0)."""
    a = np.asarray(a).astype(np.int32)
    b = np.asarray(b).astype(np.int32)
    ka = np.where(a & 0x8000, 0x8000 - (a & 0x7FFF), 0x8000 + a)  # monotone key
    kb = np.where(b & 0x8000, 0x8000 - (b & 0x7FFF), 0x8000 + b)
    d = np.abs(ka - kb)
    nan = ((a & 0x7FFF) > 0x7C00) & ((b & 0x7FFF) > 0x7C00)
    return np.where(nan, 0, d)


def compare_atlas(name, got, want, rtol=1e-3, atol=1e-4):
    """north_star tolerance for atlas texels: 1e-3 relative / 1e-4 absolute in fp32.  Returns a report dict."""
    g, w = f16(got), f16(want)
    both_nan = np.isnan(g) & np.isnan(w)
    err = np.where(both_nan, 0.0, np.abs(g - w))
    tol = atol + rtol * np.abs(w)
    bad = ~(err <= tol) & ~both_nan
    ulps = ulp16_diff(got, want)
    rep = {
        "name": name,
        "texels": int(g.size),
        "mismatched_bits": int((np.asarray(got) != np.asarray(want)).sum()),
        "max_ulp16": int(ulps.max()) if ulps.size else 0,
        "max_abs_err": float(err.max()) if err.size else 0.0,
        "out_of_tolerance": int(bad.sum()),
    }
    return rep


def compare_hit_distance(got_dd, want_dd, tol=1e-3):
    """north_star tolerance for hit distances: 1e-3 scene units (+ the fp16 storage step at that magnitude)."""
    g, w = f16(got_dd[..., 3]), f16(want_dd[..., 3])
    err = np.abs(g - w)
    return {
        "rays": int(g.size),
        "mismatched_bits": int((got_dd[..., 3] != want_dd[..., 3]).sum()),
        "max_abs_err": float(err.max()),
        "frac_within_tol": float((err <= tol).mean()),
    }
