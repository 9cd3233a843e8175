"""north_star parity tolerances and the comparisons that apply them."""
from __future__ import annotations

import math

import numpy as np

TOL_TRANSLATION_PX = 0.05
TOL_ROTATION_DEG = 0.01
TOL_SCALE = 1e-4
TOL_PIXEL = {"bilinear": 1e-3, "bicubic": 2e-3}


def decompose(m):
    m = np.asarray(m, dtype=np.float64)
    m = m / m[2, 2]
    return m[0, 2], m[1, 2], math.degrees(math.atan2(m[1, 0], m[0, 0])), math.hypot(m[0, 0], m[1, 0]), m[2, 0], m[2, 1]


def transform_delta(a, b):
    """(max |dt| px, |drot| deg, |dscale|, max |dperspective| * 1000 px) between two 3x3."""
    pa, pb = decompose(a), decompose(b)
    return (max(abs(pa[0] - pb[0]), abs(pa[1] - pb[1])), abs(pa[2] - pb[2]), abs(pa[3] - pb[3]),
            max(abs(pa[4] - pb[4]), abs(pa[5] - pb[5])) * 1000.0)


def assert_transform_close(a, b, what="", scale_px=1.0):
    dt, dr, ds, dp = transform_delta(a, b)
    assert dt <= TOL_TRANSLATION_PX * scale_px, f"{what}: translation differs by {dt} px"
    assert dr <= TOL_ROTATION_DEG, f"{what}: rotation differs by {dr} deg"
    assert ds <= TOL_SCALE, f"{what}: scale differs by {ds}"
    assert dp <= TOL_TRANSLATION_PX, f"{what}: perspective terms move a point 1000 px out by {dp} px"
    return dt, dr, ds


def compare_nested(a, b, path="", atol=2e-5, rtol=2e-5):
    """Key-set equality + numeric tolerance, in the spirit of the reference's
    scripts/compare_refactor_behavior.py:196-217."""
    if isinstance(a, dict):
        assert isinstance(b, dict), path
        assert set(a) == set(b), f"{path}: keys differ: {sorted(set(a) ^ set(b))}"
        for k in a:
            compare_nested(a[k], b[k], f"{path}.{k}", atol, rtol)
    elif isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            compare_nested(x, y, f"{path}[{i}]", atol, rtol)
    elif isinstance(a, bool) or a is None or isinstance(a, str):
        assert a == b, f"{path}: {a!r} != {b!r}"
    elif isinstance(a, (int, float)):
        assert isinstance(b, (int, float)) and not isinstance(b, bool), path
        assert math.isclose(a, b, rel_tol=rtol, abs_tol=atol), f"{path}: {a} != {b}"
    else:
        raise AssertionError(f"{path}: unexpected type {type(a)}")
