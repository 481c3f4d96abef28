"""the reference's UNMODIFIED Frame class (imported from /root/reference, which only this build container
has) constructed on top of pyorbslam_b200.ORBextractor objects whose extract_arrays() replays what the CUDA path returned on a B200
(tests/golden/gpu_extract_small.npz, recorded by tests/golden/record_gpu_extract.py) -- everything above that call is the shipped
code: operator_kd's 6-tuple list, the getters through the C ABI, `cv2.KeyPoint(*kp)` in Frame.ExtractORB (Frame.py:114-121), then the
reference's own compute_stereo_matches.  mvuRight / mvDepth must equal the golden produced by the reference Frame on the reference
extractor.  Skipped where the reference tree is absent (the GPU box)."""
import os
import sys

import numpy as np
import pytest

import oracle as O
from pyorbslam_b200 import ORBextractor
from pyorbslam_b200.synthetic import make_stereo_pair, pair_digest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "Frame.py")), reason="reference tree not present")


class _ReplayExtractor(ORBextractor):
    """The shipped class; only the call into the CUDA library is replaced by the arrays that call returned on the GPU box, and the
    pyramid (plain arrays, bit-identical to the oracle's by the -m gpu tests) comes from the oracle."""

    def __init__(self, params, kps, desc, pyramid):
        super().__init__(*params)
        self._rec = (kps, desc)
        self._pyr = pyramid

    def extract_arrays(self, image):
        return self._rec[0].copy(), self._rec[1].copy()

    def GetImagePyramid(self):
        return [p.copy() for p in self._pyr]


def test_unmodified_reference_frame_runs_on_our_extractor_outputs(golden_dir):
    cv2 = pytest.importorskip("cv2")
    sys.path.insert(0, REF)
    try:
        from Frame import Frame          # the reference's own class, unchanged
    finally:
        sys.path.remove(REF)
    g = np.load(os.path.join(golden_dir, "stereo_small.npz"))
    rec = np.load(os.path.join(golden_dir, "gpu_extract_small.npz"))
    L, R = make_stereo_pair(int(g["idx"]), int(g["H"]), int(g["W"]))
    assert pair_digest(L, R) == str(rec["image_digest"]) == str(g["image_digest"])
    assert int(rec["kernel_launches"]) >= 20                                   # the record really came from the CUDA path
    p = g["params"]
    params = (int(p[0]), float(p[1]), int(p[2]), int(p[3]), int(p[4]))
    # the recorded CUDA outputs are the reference extractor's outputs bit for bit
    assert np.array_equal(rec["kpsL"].view(np.uint32), g["kpsL"].view(np.uint32)) and np.array_equal(rec["descL"], g["descL"])
    assert np.array_equal(rec["kpsR"].view(np.uint32), g["kpsR"].view(np.uint32)) and np.array_equal(rec["descR"], g["descR"])
    oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
    oL.extract_arrays(L)
    oR.extract_arrays(R)
    eL = _ReplayExtractor(params, rec["kpsL"], rec["descL"], oL.GetImagePyramid())
    eR = _ReplayExtractor(params, rec["kpsR"], rec["descR"], oR.GetImagePyramid())
    fx, fy, cx, cy, mbf = (float(g[k]) for k in ("fx", "fy", "cx", "cy", "mbf"))
    W, H = int(g["W"]), int(g["H"])
    mK = np.eye(3, dtype=np.float32)
    mK[0, 0], mK[1, 1], mK[0, 2], mK[1, 2] = fx, fy, cx, cy
    frame_args = [fx, fy, cx, cy, 1.0 / fx, 1.0 / fy, 64.0 / W, 48.0 / H, 0.0, float(W), 0.0, float(H), 48, 64]      # Tracking.py:97-109
    f = Frame(L, R, 0.0, eL, eR, None, mK, np.zeros((4, 1), np.float32), mbf, mbf * 35 / fx, frame_args)
    assert f.N == len(rec["kpsL"]) and all(isinstance(k, cv2.KeyPoint) for k in f.mvKeys)
    kk = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in f.mvKeys], np.float32)
    assert np.array_equal(kk.view(np.uint32), g["kpsL"].view(np.uint32))         # cv2.KeyPoint(*our tuple) holds exactly the reference's fields
    assert f.mDescriptors.dtype == np.uint8 and np.array_equal(f.mDescriptors, g["descL"])
    assert np.array_equal(np.array([float(v) for v in f.mvuRight]), g["uRight"])
    assert np.array_equal(np.array([float(v) for v in f.mvDepth]), g["depth"])
    assert f.mvScaleFactors == oL.sf.tolist() and f.mnScaleLevels == params[2]
    assert sum(len(c) for col in f.mGrid for c in col) > 0.9 * f.N             # assign_features_to_grid consumed mvKeysUn
