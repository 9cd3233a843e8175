"""Pins oracle/dis_ref.c (plain-C DIS restatement) against live cv2 and the committed goldens."""
import os

import numpy as np
import pytest

from oracle import dis_ref
from tests import cases
from tests.conftest import GOLDEN_DIR


def _cv2_dis(cv2, vr=5, finest=2):
    d = cv2.DISOpticalFlow.create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
    d.setFinestScale(finest)
    d.setPatchSize(8)
    d.setPatchStride(4)
    d.setUseSpatialPropagation(True)
    d.setVariationalRefinementIterations(vr)
    return d


@pytest.mark.parametrize("case", cases.DIS_CASES, ids=[c["name"] for c in cases.DIS_CASES])
def test_dis_oracle_matches_reference_golden(case):
    """Golden = output of the reference's own _create_flow_backend('DIS').calc (cv2 4.13.0.92)."""
    gold = np.load(os.path.join(GOLDEN_DIR, f"dis_{case['name']}.npz"))
    prev, curr = cases.make_gray_pair(case)
    flow = dis_ref.calc(prev, curr)
    assert np.array_equal(flow[::8, ::8], gold["grid"])
    if case["store"] == "full":
        assert np.array_equal(flow, gold["flow"])


@pytest.mark.parametrize("size", [(320, 180), (640, 360), (333, 187)])
@pytest.mark.parametrize("vr", [0, 5])
def test_dis_oracle_bit_exact_vs_live_cv2(size, vr):
    cv2 = pytest.importorskip("cv2")
    w, h = size
    prev, curr = cases.make_gray_pair(dict(w=w, h=h, seed=w + vr, amount=2.0))
    ref = _cv2_dis(cv2, vr).calc(prev, curr, None)
    mine = dis_ref.calc(prev, curr, dis_ref.default_params(vr_iter=vr))
    assert np.array_equal(ref, mine)


def test_dis_oracle_white_noise():
    """Noise maximises the number of discrete decisions (candidate picks, early stops)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    i0 = rng.integers(0, 256, (270, 480), dtype=np.uint8)
    i1 = np.roll(i0, (2, 3), (0, 1))
    assert np.array_equal(_cv2_dis(cv2).calc(i0, i1, None), dis_ref.calc(i0, i1))


@pytest.mark.parametrize("size", [(60, 33), (61, 34), (120, 67)])
def test_variational_refinement_bit_exact(size):
    cv2 = pytest.importorskip("cv2")
    w, h = size
    prev, curr = cases.make_gray_pair(dict(w=w, h=h, seed=3, amount=0.2))
    rng = np.random.default_rng(0)
    u = cv2.GaussianBlur(rng.normal(0, 0.3, (h, w)).astype(np.float32), (0, 0), 3)
    v = cv2.GaussianBlur(rng.normal(0, 0.3, (h, w)).astype(np.float32), (0, 0), 3)
    vr = cv2.VariationalRefinement.create()
    vr.setEpsilon(0.01)
    ru, rv = u.copy(), v.copy()
    vr.calcUV(prev, curr, ru, rv)
    mu, mv = dis_ref.variational_refinement(prev, curr, u, v)
    assert np.array_equal(ru, mu) and np.array_equal(rv, mv)


@pytest.mark.parametrize("shape", [((120, 67), (240, 135)), ((30, 16), (60, 33))])
def test_flow_upsampling_paths(shape):
    cv2 = pytest.importorskip("cv2")
    (sw, sh), (dw, dh) = shape
    rng = np.random.default_rng(1)
    a = rng.normal(0, 3, (sh, sw)).astype(np.float32)
    assert np.array_equal(cv2.resize(a, (dw, dh)), dis_ref.resize_linear_f32(a, (dw, dh), mode=2))
    a2 = rng.normal(0, 3, (sh, sw, 2)).astype(np.float32)
    assert np.array_equal(cv2.resize(a2, (dw * 2, dh * 2)), dis_ref.resize_linear_f32(a2, (dw * 2, dh * 2), mode=0))


SMALL_SIZES = [(73, 45), (121, 73), (90, 50), (40, 24), (30, 100), (64, 48), (20, 9), (12, 12), (89, 33)]


@pytest.mark.parametrize("size", SMALL_SIZES, ids=[f"{w}x{h}" for w, h in SMALL_SIZES])
def test_dis_oracle_small_frames_follow_the_backend_state(size):
    """Frames too small for finest scale 2 (the reference's own checks run Flow on 73x45 and 121x73:
    scripts/compare_refactor_behavior.py:220, scripts/check_crop_aspect_ratio.py:165): cv2 selects the levels
    itself, computes down to full resolution and leaves the new finest scale on the object, so the second
    pair of a clip can run on fewer levels than the first (90x50, 40x24, 30x100).  One object, three calls."""
    cv2 = pytest.importorskip("cv2")
    w, h = size
    prev, curr = cases.make_gray_pair(dict(w=w, h=h, seed=w + h, amount=2.0))
    ref, mine = _cv2_dis(cv2), dis_ref.Backend()
    for a, b in ((prev, curr), (curr, prev), (prev, curr)):
        assert np.array_equal(ref.calc(a, b, None), mine.calc(a, b))
    assert ref.getFinestScale() == mine.params.finest_scale


def test_dis_oracle_scale_selection_table():
    assert dis_ref.select_scales(540, 960) == (2, 5) and dis_ref.select_scales(73, 121) == (2, 2)
    assert dis_ref.select_scales(45, 73) == (0, 1) and dis_ref.select_scales(45, 73, 0) == (0, 1)
    assert dis_ref.select_scales(50, 90) == (0, 2) and dis_ref.select_scales(50, 90, 0) == (0, 1)
    assert dis_ref.select_scales(100, 30) == (0, 0) and dis_ref.select_scales(100, 30, 0) == (0, 1)
    assert dis_ref.select_scales(10, 10) is None and dis_ref.select_scales(7, 200) is None
    with pytest.raises(ValueError, match="smaller than one patch"):  # 100x30 -> levels 2..0, level 2 is 25x7
        dis_ref.calc(np.zeros((30, 100), np.uint8), np.zeros((30, 100), np.uint8))
