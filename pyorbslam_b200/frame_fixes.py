"""SURVEY.md 8(f) rank 3 -- the correctness fix of `Frame.undistort_keypoints` (reference Frame.py:293-322).

The reference's method only works for undistorted cameras (KITTI): with k1 != 0 it reads an undefined name (`mvKeys` instead of
`self.mvKeys`, Frame.py:298-300) and, were that fixed, returns the list instead of assigning `self.mvKeysUn` (Frame.py:322), so
`Frame.__init__` (Frame.py:64) would leave `mvKeysUn` unset.  `install(Frame, fix_undistort=True)` replaces it with what the code
plainly intends -- upstream ORB-SLAM2's Frame::UndistortKeyPoints: `cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK)`
on the keypoint coordinates, every other KeyPoint field kept.  Opt-in: with the default install() Frame.py behaves exactly as shipped.

The numbers are restated here (NumPy float64, the arithmetic of OpenCV's iterative undistortPoints: 5 fixed-point iterations of the
inverse Brown-Conrady model with k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]], result re-projected with P = mK and rounded to float32 like a
CV_32FC2 destination) and pinned to cv2.undistortPoints by tests/test_frame_fixes.py; they are a few thousand flops per frame,
host work by design (the hot path's kernels are elsewhere)."""
import numpy as np


def undistort_points(xy, K, dist, iters=5):
    """xy: [n, 2] pixel coordinates; K: 3x3 camera matrix; dist: 4, 5, 8 or 12 distortion coefficients.  -> float32 [n, 2]."""
    xy = np.asarray(xy, np.float32).reshape(-1, 2).astype(np.float64)     # CV_32FC2 source, double arithmetic
    K = np.asarray(K, np.float64).reshape(3, 3)
    k = np.zeros(14, np.float64)
    d = np.asarray(dist, np.float64).reshape(-1)
    if d.size not in (4, 5, 8, 12, 14):
        raise ValueError("distortion coefficients must have 4, 5, 8, 12 or 14 entries")
    k[:d.size] = d
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    x = (xy[:, 0] - cx) / fx
    y = (xy[:, 1] - cy) / fy
    x0, y0 = x.copy(), y.copy()
    if np.any(k[12:14] != 0):
        raise ValueError("tilted sensor coefficients (tauX, tauY) are not supported")
    frozen = np.zeros(len(x), bool)          # points whose icdist went negative: OpenCV falls back to the plain normalisation and stops
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
        neg = (icdist < 0) & ~frozen
        dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
        dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
        nx, ny = (x0 - dx) * icdist, (y0 - dy) * icdist
        upd = ~frozen & ~neg
        x = np.where(upd, nx, np.where(neg, x0, x))
        y = np.where(upd, ny, np.where(neg, y0, y))
        frozen |= neg
    out = np.stack([x * fx + cx, y * fy + cy], 1)        # P = K, R = identity
    return out.astype(np.float32)


def undistort_keypoints(self):
    """Drop-in body for Frame.undistort_keypoints: sets self.mvKeysUn (Frame.py:293-322 as intended)."""
    if self.mDistCoef[0][0] == 0:                        # Frame.py:295-297, unchanged
        self.mvKeysUn = self.mvKeys
        return
    pts = np.array([[kp.pt[0], kp.pt[1]] for kp in self.mvKeys], np.float32).reshape(-1, 2)
    und = undistort_points(pts, self.mK, np.asarray(self.mDistCoef).reshape(-1)) if len(pts) else pts
    out = []
    for i, kp in enumerate(self.mvKeys):                 # a new KeyPoint of the same class, only the position changes (Frame.py:310-320)
        out.append(type(kp)(x=float(und[i, 0]), y=float(und[i, 1]), size=kp.size, angle=kp.angle, response=kp.response,
                            octave=kp.octave, class_id=kp.class_id))
    self.mvKeysUn = out
