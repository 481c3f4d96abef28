"""Pins the oracle's OpenCV-owned primitives bit-for-bit against cv2 (the reference's third-party
dependency, OpenCV 4.x -- here cv2 4.13.0) and its sinf/cosf against the host libm.
Reference call sites: ORBextractor.cpp:1120 (resize), :1122-1128 (copyMakeBorder), :808-814 (FAST),
:1085 (GaussianBlur), :103 (fastAtan2), :113 (cosf/sinf)."""
import numpy as np
import pytest

import oracle as O

cv2 = pytest.importorskip("cv2")


def _rng():
    return np.random.default_rng(1234)


def test_border_reflect101():
    img = _rng().integers(0, 256, (41, 57), dtype=np.uint8)
    assert np.array_equal(O.border101(img), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
    tiny = _rng().integers(0, 256, (5, 9), dtype=np.uint8)   # border wider than the image: multiple reflections
    assert np.array_equal(O.border101(tiny), cv2.copyMakeBorder(tiny, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("W,H", [(1241, 376), (1226, 370), (2560, 1440), (640, 480), (97, 131)])
def test_resize_chain_matches_cv2(W, H):
    img = _rng().integers(0, 256, (H, W), dtype=np.uint8)
    cur, s = img, np.float32(1.0)
    for _ in range(1, 8):
        s = np.float32(np.float64(s) * np.float64(np.float32(1.2)))
        inv = np.float32(1.0) / s
        dw, dh = int(np.rint(np.float32(W) * inv)), int(np.rint(np.float32(H) * inv))
        if dw < 4 or dh < 4:
            break
        ref = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(O.resize(cur, dw, dh), ref)
        cur = ref


@pytest.mark.parametrize("scale", [1.01, 1.05, 1.1, 1.3, 1.5, 2.0])
def test_resize_other_scales(scale):
    img = _rng().integers(0, 256, (376, 1241), dtype=np.uint8)
    dw, dh = int(round(1241 / scale)), int(round(376 / scale))
    assert np.array_equal(O.resize(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_gaussian_blur():
    for shape in [(97, 131), (376, 1241), (8, 9)]:
        img = _rng().integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(O.blur7(img), cv2.GaussianBlur(img, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101))


def _cv_fast(im, t):
    det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    return np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(im, None)], np.int32).reshape(-1, 3)


def test_fast_matches_cv2_including_ties_and_order():
    rng = _rng()
    for trial in range(40):
        h, w = int(rng.integers(7, 60)), int(rng.integers(7, 70))
        if trial % 2 == 0:
            im = rng.integers(0, 256, (h, w), dtype=np.uint8)
        else:
            im = (rng.integers(0, 4, (h, w)) * 60).astype(np.uint8)   # many equal scores -> NMS ties
        for t in (20, 7, 1, 0):
            assert np.array_equal(O.fast(im, t), _cv_fast(im, t)), (trial, t)
    assert len(O.fast(np.zeros((6, 30), np.uint8), 7)) == 0       # smaller than the 7x7 support


def test_fast_atan2():
    rng = _rng()
    y = rng.integers(-200000, 200000, 4000).astype(np.float32)
    x = rng.integers(-200000, 200000, 4000).astype(np.float32)
    y[:10] = 0
    x[:5] = 0
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(O.atan2_deg(y, x).view(np.uint32), ref.view(np.uint32))


def test_sincosf_exhaustive_vs_host_libm():
    # every float in [0, 2*pi] (1.09e9 values), both functions
    assert O.lib().orbo_sincos_exhaustive_mismatches() == 0
