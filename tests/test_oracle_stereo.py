"""Stereo restatements (C `orbo_stereo`, Python `stereo_py`) vs vectors produced by the reference's own
Frame.compute_stereo_matches (Frame.py:161-279), see tests/golden/make_golden.py."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import stereo_py
from pyorbslam_b200.synthetic import make_kitti_like_pair, make_stereo_pair, pair_digest


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    gen = make_kitti_like_pair if "kitti_like" in name else make_stereo_pair
    L, R = gen(int(g["idx"]), int(g["H"]), int(g["W"]))
    assert pair_digest(L, R) == str(g["image_digest"]), "synthetic generator drifted from the fixture"
    p = g["params"]
    params = (int(p[0]), float(p[1]), int(p[2]), int(p[3]), int(p[4]))
    return g, L, R, params


@pytest.mark.parametrize("name", ["stereo_kitti_shape.npz", "stereo_small.npz", "stereo_kitti_like.npz"])
def test_c_stereo_bit_exact_vs_reference_frame(golden_dir, name):
    g, L, R, params = _load(golden_dir, name)
    eL, eR = O.OracleExtractor(*params), O.OracleExtractor(*params)
    kL, dL = eL.extract_arrays(L)
    kR, dR = eR.extract_arrays(R)
    # the extractor part must already agree with what the reference Frame saw
    assert np.array_equal(kL.view(np.uint32), g["kpsL"].view(np.uint32)) and np.array_equal(dL, g["descL"])
    assert np.array_equal(kR.view(np.uint32), g["kpsR"].view(np.uint32)) and np.array_equal(dR, g["descR"])
    uR, dep, _, _ = O.stereo(kL[:, [0, 1, 5]], dL, kR[:, [0, 1, 5]], dR, eL.sf, eL.isf,
                             eL.GetImagePyramid(), eR.GetImagePyramid(), float(g["mbf"]), float(g["fx"]))
    gu, gd = g["uRight"], g["depth"]
    assert np.array_equal(uR >= 0, gu >= 0)                       # match decisions: exact
    assert np.array_equal(uR.astype(np.float64), gu)              # values: bit-exact (float32 both sides)
    assert np.array_equal(dep.astype(np.float64), gd)
    assert (gu >= 0).sum() > 100


def test_python_restatement_vs_reference_frame(golden_dir):
    g, L, R, params = _load(golden_dir, "stereo_small.npz")
    eL, eR = O.OracleExtractor(*params), O.OracleExtractor(*params)
    tL, dL = eL.operator_kd(L)
    tR, dR = eR.operator_kd(R)
    keysL = [(k[0], k[1], k[5]) for k in tL]
    keysR = [(k[0], k[1], k[5]) for k in tR]
    u, d = stereo_py.stereo_matches(keysL, dL, keysR, dR, eL.GetScaleFactors(), eL.GetInverseScaleFactors(),
                                    eL.GetImagePyramid(), eR.GetImagePyramid(), float(g["mbf"]), np.float32(g["fx"]))
    assert np.array_equal(np.array([float(v) for v in u]), g["uRight"])
    assert np.array_equal(np.array([float(v) for v in d]), g["depth"])


def test_stereo_no_right_keypoints_and_no_left():
    e = O.OracleExtractor(300, 1.2, 4, 20, 7)
    L, R = make_stereo_pair(9, 160, 320)
    kL, dL = e.extract_arrays(L)
    pyr = e.GetImagePyramid()
    empty_k, empty_d = np.zeros((0, 3), np.float32), np.zeros((0, 32), np.uint8)
    uR, dep, bi, _ = O.stereo(kL[:, [0, 1, 5]], dL, empty_k, empty_d, e.sf, e.isf, pyr, pyr, 100.0, 300.0)
    assert (uR == -1).all() and (dep == -1).all() and (bi == -1).all()
    uR, dep, _, _ = O.stereo(empty_k, empty_d, kL[:, [0, 1, 5]], dL, e.sf, e.isf, pyr, pyr, 100.0, 300.0)
    assert len(uR) == 0


def test_stereo_identical_views_match_themselves():
    # left == right: every keypoint's Hamming winner is itself (distance 0); the SAD minimum sits at incR=0, so the
    # sub-pixel disparity is within one level-pixel, and exact zeros take the 0.01 clamp (Frame.py:273-275)
    e = O.OracleExtractor(300, 1.2, 4, 20, 7)
    L, _ = make_stereo_pair(9, 160, 320)
    k, d = e.extract_arrays(L)
    pyr = e.GetImagePyramid()
    uR, dep, bi, bd = O.stereo(k[:, [0, 1, 5]], d, k[:, [0, 1, 5]], d, e.sf, e.isf, pyr, pyr, 100.0, 300.0)
    assert (bd == 0).all() and np.array_equal(bi, np.arange(len(k)))
    ok = uR >= 0
    assert ok.sum() > 0.3 * len(k)
    disp = k[ok, 0] - uR[ok]
    assert (disp > 0).all() and (disp <= e.sf[k[ok, 5].astype(int)] + 0.011).all()
    assert np.allclose(dep[ok], np.float32(100.0) / disp, rtol=1e-4)


def test_median_cull_restatement_semantics():
    # upstream ORB-SLAM2: sorted SAD minima, median = element [n // 2], thDist = 1.5f * 1.4f * median, drop dist >= thDist
    sad = np.array([-1, 10, 40, 20, -1, 30, 100, 63, 62], np.int32)
    u = np.arange(9, dtype=np.float32) + 1
    d = u * 2
    cu, cd = O.median_cull(u, d, sad)
    # accepted = [10, 20, 30, 40, 62, 63, 100] -> median = 40 -> thDist = 2.1f * 40 = 84 -> only 100 is dropped
    assert cu.tolist() == [1, 2, 3, 4, 5, 6, -1, 8, 9] and cd[6] == -1
    cu, _ = O.median_cull(u, d, np.full(9, -1, np.int32))
    assert np.array_equal(cu, u)
