"""GPU replacement for `Frame.compute_stereo_matches` (reference Frame.py:161-279).

`install(Frame)` assigns it to the class attribute, exactly the hook the reference's README names
(README.md:13); Frame.py itself stays byte-identical.  When the frame's descriptors are the arrays our
extractor objects handed out last (the normal `Frame.__init__` order, Frame.py:48-65) the matcher runs on
the device-resident keypoints / descriptors / pyramids; otherwise it uploads the frame's own host data."""
import ctypes as C

import numpy as np

from . import _lib
from .extractor import ORBextractor


def _to_lists(uR, dep, mbf=None, uL=None):
    # reference: python int -1 where unmatched (Frame.py:163-164); matched entries are the NumPy float32 SCALARS that
    # `bestuR = mvScaleFactors[...] * (...)` and `mbf / disparity` leave under NumPy >= 2 (Frame.py:269,277-278) -- the
    # element type decides the precision of downstream expressions such as abs(ur - mvuRight[i]) (ORBMatcher.py:356-358).
    u = list(uR)
    d = list(dep)
    m = uR < 0
    if m.any():
        for i in np.nonzero(m)[0].tolist():
            u[i] = -1
            d[i] = -1
    if mbf is not None and uL is not None:
        # the disparity <= 0 branch (Frame.py:273-275) leaves Python floats built from the literal 0.01; the kernel stores their
        # float32 roundings, recognisable by the depth value mbf / 0.01
        z = np.nonzero(dep == np.float32(float(mbf) / 0.01))[0]
        for i in z.tolist():
            u[i] = float(uL(i)) - 0.01
            d[i] = float(mbf) / 0.01
    return u, d


MEDIAN_CULL = 1      # B200ORB_STEREO_MEDIAN_CULL   -- upstream ORB-SLAM2 behaviour, NOT the reference's
DENSE_PYRAMID = 2    # B200ORB_STEREO_DENSE_PYRAMID -- true level images instead of the reference's sheared view (F6)
OPTIONS = {"flags": 0}   # what the installed Frame.compute_stereo_matches uses; 0 = exactly the reference


def stereo_resident(extL, extR, mbf, fx, flags=0, with_sad=False):
    n = extL._last_n
    uR = np.empty(max(n, 0), np.float32)
    dep = np.empty(max(n, 0), np.float32)
    mi = np.empty(max(n, 0), np.int32)
    sad = np.empty(max(n, 0), np.int32)
    if n > 0:
        _lib.check(_lib.lib().b200orb_stereo_ex(extL._h, extR._h, float(mbf), float(np.float32(fx)), int(flags),
                                                uR.ctypes.data, dep.ctypes.data, mi.ctypes.data, sad.ctypes.data))
    if with_sad:
        return uR, dep, mi, sad
    return uR, dep, mi


def stereo_host(kpsL, descL, kpsR, descR, sf, isf, pyrL, pyrR, mbf, fx, device=0):
    """Array form on caller data: kps* float32[n,3] = (x, y, octave); pyr* = GetImagePyramid() views."""
    kpsL = np.ascontiguousarray(kpsL, np.float32).reshape(-1, 3)
    kpsR = np.ascontiguousarray(kpsR, np.float32).reshape(-1, 3)
    descL = np.ascontiguousarray(descL, np.uint8).reshape(-1, 32) if len(kpsL) else np.zeros((0, 32), np.uint8)
    descR = np.ascontiguousarray(descR, np.uint8).reshape(-1, 32) if len(kpsR) else np.zeros((0, 32), np.uint8)
    sf = np.ascontiguousarray(sf, np.float32)
    isf = np.ascontiguousarray(isf, np.float32)
    L = len(sf)
    pl = [np.ascontiguousarray(p, np.uint8) for p in pyrL]
    pr = [np.ascontiguousarray(p, np.uint8) for p in pyrR]
    PL = (C.c_void_p * L)(*[p.ctypes.data for p in pl])
    PR = (C.c_void_p * L)(*[p.ctypes.data for p in pr])
    lw = np.array([p.shape[1] for p in pl], np.int32)
    lh = np.array([p.shape[0] for p in pl], np.int32)
    n = len(kpsL)
    uR = np.empty(n, np.float32)
    dep = np.empty(n, np.float32)
    mi = np.empty(n, np.int32)
    _lib.check(_lib.lib().b200orb_stereo_host(int(device), n, kpsL.ctypes.data, descL.ctypes.data, len(kpsR), kpsR.ctypes.data,
                                              descR.ctypes.data, L, sf.ctypes.data, isf.ctypes.data, C.cast(PL, C.c_void_p),
                                              C.cast(PR, C.c_void_p), lw.ctypes.data, lh.ctypes.data, float(mbf),
                                              float(np.float32(fx)), uR.ctypes.data, dep.ctypes.data, mi.ctypes.data))
    return uR, dep, mi


def compute_stereo_matches(self):
    """Drop-in body for Frame.compute_stereo_matches: fills self.mvuRight / self.mvDepth (length N, -1 = no match)."""
    extL, extR = self.mpORBextractorLeft, self.mpORBextractorRight
    fx = self.mK[0][0]
    resident = (isinstance(extL, ORBextractor) and isinstance(extR, ORBextractor)
                and getattr(self, "mDescriptors", None) is extL._last_desc
                and getattr(self, "mDescriptorsRight", None) is extR._last_desc
                and extL._last_n == self.N)
    if resident:
        uR, dep, _ = stereo_resident(extL, extR, self.mbf, fx, flags=OPTIONS["flags"])
    else:
        kL = np.array([[k.pt[0], k.pt[1], k.octave] for k in self.mvKeys], np.float32).reshape(-1, 3)
        kR = np.array([[k.pt[0], k.pt[1], k.octave] for k in self.mvKeysRight], np.float32).reshape(-1, 3)
        uR, dep, _ = stereo_host(kL, self.mDescriptors, kR, self.mDescriptorsRight, self.mvScaleFactors, self.mvInvScaleFactors,
                                 self.mvImagePyramidLeft, self.mvImagePyramidRight, self.mbf, fx,
                                 device=getattr(extL, "_device", 0))
    self.mvuRight, self.mvDepth = _to_lists(uR, dep, self.mbf, lambda i: self.mvKeys[i].pt[0])


def install(frame_cls, median_cull=False, dense_pyramid=False, fix_undistort=False):
    """Frame.compute_stereo_matches = the GPU matcher.  Returns the original method (to restore / compare).
    median_cull / dense_pyramid switch on upstream-ORB-SLAM2 behaviour the reference does not have (device-resident path
    only); fix_undistort replaces Frame.undistort_keypoints, which is broken for distorted cameras (Frame.py:298-322:
    undefined name, result never assigned), by what it intends (frame_fixes.py).  Leave all three off for the reference's
    behaviour exactly as shipped."""
    OPTIONS["flags"] = (MEDIAN_CULL if median_cull else 0) | (DENSE_PYRAMID if dense_pyramid else 0)
    original = frame_cls.compute_stereo_matches
    frame_cls.compute_stereo_matches = compute_stereo_matches
    if fix_undistort:
        from .frame_fixes import undistort_keypoints
        frame_cls.undistort_keypoints = undistort_keypoints
    return original
