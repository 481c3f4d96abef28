"""the opt-in fix of Frame.undistort_keypoints (reference Frame.py:293-322) -- the restated OpenCV
arithmetic against cv2.undistortPoints itself (the call the reference's code makes), and the method on a Frame-shaped object."""
import types

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from pyorbslam_b200.frame_fixes import undistort_keypoints, undistort_points   # noqa: E402
from pyorbslam_b200.stereo import install                                       # noqa: E402

K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]], np.float32)
DISTS = {
    "k1k2p1p2": [-0.28, 0.07, 0.0002, -0.0003],
    "with_k3": [-0.35, 0.15, 0.001, -0.0007, -0.03],
    "rational": [0.1, -0.05, 0.0005, 0.0002, 0.01, 0.2, -0.1, 0.02],
    "thin_prism": [-0.2, 0.05, 0.001, 0.001, 0.0, 0.0, 0.0, 0.0, 0.001, -0.0005, 0.0007, 0.0002],
    "strong": [-0.6, 0.4, 0.0, 0.0],
}


@pytest.mark.parametrize("name", sorted(DISTS))
def test_restated_undistort_points_equals_cv2(name):
    rng = np.random.default_rng(4)
    pts = np.stack([rng.uniform(0, 1241, 3000), rng.uniform(0, 376, 3000)], 1).astype(np.float32)
    dist = np.array(DISTS[name], np.float32).reshape(-1, 1)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, None, K).reshape(-1, 2)      # Frame.py:306
    got = undistort_points(pts, K, dist)
    assert got.dtype == np.float32 and got.shape == ref.shape
    # float32 output of a double computation: identical up to the last float32 bit of a ~1e3 px coordinate
    assert np.abs(got.astype(np.float64) - ref).max() <= 1.3e-4
    assert (got == ref).mean() > 0.9


def test_fixed_method_assigns_mvkeysun_and_keeps_the_other_fields():
    class Frame:                                  # the attributes Frame.undistort_keypoints reads (Frame.py:293-322)
        def compute_stereo_matches(self):
            pass

        def undistort_keypoints(self):            # the reference's body fails like this for k1 != 0 (undefined name, Frame.py:298)
            raise NameError("name 'mvKeys' is not defined")
    install(Frame, fix_undistort=True)
    f = Frame()
    f.mK = K
    f.mDistCoef = np.array(DISTS["with_k3"], np.float32).reshape(-1, 1)
    f.mvKeys = [cv2.KeyPoint(100.0 + 37 * i, 50.0 + 11 * i, 31.0, 12.5 * i, 40.0 + i, i % 8) for i in range(20)]
    f.undistort_keypoints()
    assert len(f.mvKeysUn) == 20 and all(isinstance(k, cv2.KeyPoint) for k in f.mvKeysUn)
    pts = np.array([k.pt for k in f.mvKeys], np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, f.mDistCoef, None, None, K).reshape(-1, 2)
    for a, b, r in zip(f.mvKeys, f.mvKeysUn, ref):
        assert abs(b.pt[0] - r[0]) < 2e-4 and abs(b.pt[1] - r[1]) < 2e-4
        assert (b.size, b.angle, b.response, b.octave, b.class_id) == (a.size, a.angle, a.response, a.octave, a.class_id)
        assert b.pt != a.pt
    # k1 == 0 (every KITTI sequence): the reference's own branch, mvKeysUn IS mvKeys (Frame.py:295-297)
    f.mDistCoef = np.zeros((4, 1), np.float32)
    f.undistort_keypoints()
    assert f.mvKeysUn is f.mvKeys
    # the default install leaves Frame.undistort_keypoints alone
    class Frame2(Frame):
        pass
    before = Frame2.undistort_keypoints
    install(Frame2)
    assert Frame2.undistort_keypoints is before


def test_empty_keypoint_list():
    f = types.SimpleNamespace(mK=K, mDistCoef=np.array(DISTS["k1k2p1p2"], np.float32).reshape(-1, 1), mvKeys=[])
    undistort_keypoints(f)
    assert f.mvKeysUn == []
