"""Runs on a GPU box: records what pyorbslam_b200.ORBextractor (the CUDA path, through the C ABI) returns for the stereo_small golden
pair, so that the build container -- which has the reference tree but no GPU -- can feed exactly those arrays to the reference's
UNMODIFIED Frame class (tests/test_reference_frame_on_gpu_outputs.py).

Usage (GPU box): python tests/golden/record_gpu_extract.py gpurun_out/gpu_extract_small.npz   -> copy to tests/golden/"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyorbslam_b200 import ORBextractor, _lib  # noqa: E402
from pyorbslam_b200.synthetic import make_stereo_pair, pair_digest  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "stereo_small.npz"))
L, R = make_stereo_pair(int(g["idx"]), int(g["H"]), int(g["W"]))
p = g["params"]
params = (int(p[0]), float(p[1]), int(p[2]), int(p[3]), int(p[4]))
before = _lib.kernel_launches()
eL, eR = ORBextractor(*params), ORBextractor(*params)
kL, dL = eL.extract_arrays(L)
kR, dR = eR.extract_arrays(R)
assert _lib.kernel_launches() > before, "no kernel ran"
np.savez_compressed(sys.argv[1], image_digest=pair_digest(L, R), params=np.array(params, np.float64), kpsL=kL, descL=dL, kpsR=kR, descR=dR,
                    kernel_launches=_lib.kernel_launches() - before)
print("recorded", len(kL), len(kR), "keypoints from", _lib.kernel_launches() - before, "kernel launches")
