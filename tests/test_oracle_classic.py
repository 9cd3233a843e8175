"""Pins oracle/classic_ref.c (Shi-Tomasi corners + pyramidal LK) against live cv2."""
import numpy as np
import pytest

from oracle import classic_ref as CR
from tests import cases

cv2 = pytest.importorskip("cv2")


# widths that are not a multiple of 32 exercise the scalar tail of the wheel's Sobel row filter (portrait clips work at 540x960)
@pytest.mark.parametrize("size", [(960, 540), (832, 480), (640, 360), (540, 960), (333, 187), (200, 120), (121, 73), (73, 45)])
def test_gftt_identical_and_lk_close(size):
    w, h = size
    p, c = cases.make_gray_pair(dict(w=w, h=h, seed=w + 1, amount=1.5))
    assert np.array_equal(cv2.cornerMinEigenVal(p, 21, ksize=3), CR.min_eigen(p))
    f = cv2.goodFeaturesToTrack(p, maxCorners=400, qualityLevel=0.01, minDistance=7, blockSize=21).reshape(-1, 2)
    g = CR.good_features(p)
    assert np.array_equal(f, g)  # same corners in the same order
    nxt, st, _ = cv2.calcOpticalFlowPyrLK(p, c, f.reshape(-1, 1, 2), None, winSize=(31, 31), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 50, 0.01))
    mine, ms = CR.pyr_lk(p, c, f)
    st = st.ravel()
    assert int((st != ms).sum()) == 0
    ok = st == 1
    assert float(np.abs(nxt.reshape(-1, 2)[ok] - mine[ok]).max()) <= 2e-3  # float accumulation order only


def test_lk_loses_points_that_leave_the_frame():
    w, h = 320, 180
    p, _ = cases.make_gray_pair(dict(w=w, h=h, seed=3, amount=1.0))
    c = np.roll(p, (0, 60), (0, 1))  # 60 px shift: far beyond what 3 levels of a 31-px window recover at the edge
    f = cv2.goodFeaturesToTrack(p, maxCorners=100, qualityLevel=0.01, minDistance=7, blockSize=21).reshape(-1, 2)
    nxt, st, _ = cv2.calcOpticalFlowPyrLK(p, c, f.reshape(-1, 1, 2), None, winSize=(31, 31), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 50, 0.01))
    mine, ms = CR.pyr_lk(p, c, f)
    assert int((st.ravel() != ms).sum()) <= 1


@pytest.mark.parametrize("size", [(960, 540), (540, 960), (333, 187), (121, 73), (640, 360)])
def test_lk_in_cv2_lane_order_is_bit_exact(size):
    """pyr_lk(exact=True) sums the 31x31 window terms in the lane order of cv2's SIMD code (classic_ref.c): every
    tracked position carries cv2's bits, every status flag is cv2's.  (The default order is the one the CUDA tracker
    follows; moving the kernel to this one is what would make Classic as exact as Flow.)"""
    w, h = size
    for seed, amount in ((1, 1.5), (2, 6.0)):
        p, c = cases.make_gray_pair(dict(w=w, h=h, seed=w + seed, amount=amount))
        f = cv2.goodFeaturesToTrack(p, maxCorners=400, qualityLevel=0.01, minDistance=7, blockSize=21).reshape(-1, 2)
        nxt, st, _ = cv2.calcOpticalFlowPyrLK(p, c, f.reshape(-1, 1, 2), None, winSize=(31, 31), maxLevel=3,
                                              criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 50, 0.01))
        mine, ms = CR.pyr_lk(p, c, f, exact=True)
        st = st.ravel()
        assert np.array_equal(st, ms)
        assert np.array_equal(nxt.reshape(-1, 2)[st == 1], mine[st == 1])
