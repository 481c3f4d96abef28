"""TEST INFRASTRUCTURE -- CPU oracle for the stereo ORB front-end.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``pyorbslam_b200``) never does.

ctypes front for ``oracle/liborb_oracle.so`` (built from ``orb_oracle.cpp`` by ``oracle/Makefile``), a
restatement of /root/reference/pyORBExtractor/ORBextractor.cpp and /root/reference/Frame.py:161-279.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u8p = C.POINTER(C.c_ubyte)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liborb_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "cvprims.hpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        import sys
        subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "liborb_oracle.so")], stdout=sys.stderr)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orbo_create.restype = C.c_void_p
        _LIB.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.orbo_destroy.argtypes = [C.c_void_p]
        _LIB.orbo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _LIB.orbo_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        _LIB.orbo_level_size.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
        for f in ("orbo_level_bordered", "orbo_level_blurred", "orbo_level_caster_view"):
            getattr(_LIB, f).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _LIB.orbo_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _LIB.orbo_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _LIB.orbo_sincos_exhaustive_mismatches.restype = C.c_long
        _LIB.orbo_stereo_ex.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_double, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- primitives
def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orbo_resize(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def border101(src, b=19):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((src.shape[0] + 2 * b, src.shape[1] + 2 * b), np.uint8)
    lib().orbo_border101(_p(src), src.shape[1], src.shape[0], _p(dst), b)
    return dst


def blur7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    lib().orbo_blur7(_p(src), src.shape[1], src.shape[0], _p(dst))
    return dst


def fast(img, threshold):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.empty((max(cap, 1), 3), np.int32)
    n = lib().orbo_fast(_p(img), img.shape[1], img.shape[0], int(threshold), cap, _p(out))
    return out[:n].copy()


def atan2_deg(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().orbo_atan2(_p(y), _p(x), _p(out), y.size)
    return out


def sincos(a):
    a = np.ascontiguousarray(a, np.float32)
    s = np.empty_like(a)
    c = np.empty_like(a)
    lib().orbo_sincos(_p(a), _p(s), _p(c), a.size)
    return s, c


def distribute(cand, min_x, max_x, min_y, max_y, n_want):
    cand = np.ascontiguousarray(cand, np.int32).reshape(-1, 3)
    cap = len(cand) + 8
    out = np.empty((cap, 3), np.int32)
    n = lib().orbo_distribute(_p(cand), len(cand), min_x, max_x, min_y, max_y, n_want, cap, _p(out))
    return out[:n].copy()


# ---------------------------------------------------------------- extractor
class OracleExtractor:
    """Same surface as pyORBExtractor.ORBextractor (orb_extractor.cpp:22-38), CPU oracle behind it."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        self._h = lib().orbo_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        if not self._h:
            raise ValueError("bad extractor parameters")
        self.nlevels = int(nlevels)
        self.nfeatures = int(nfeatures)
        self._scale = float(np.float32(scaleFactor))
        L = self.nlevels
        self.sf = np.empty(L, np.float32)
        self.isf = np.empty(L, np.float32)
        self.sig2 = np.empty(L, np.float32)
        self.isig2 = np.empty(L, np.float32)
        self.quota = np.empty(L, np.int32)
        self.umax = np.empty(16, np.int32)
        lib().orbo_tables(self._h, _p(self.sf), _p(self.isf), _p(self.sig2), _p(self.isig2), _p(self.quota), _p(self.umax))
        self._have = False

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orbo_destroy(self._h)
            self._h = None

    def extract_arrays(self, image):
        image = np.ascontiguousarray(image, np.uint8)
        assert image.ndim == 2
        cap = self.nfeatures + 2 * self.nlevels + 64 * self.nlevels
        kps = np.empty((cap, 6), np.float32)
        desc = np.empty((cap, 32), np.uint8)
        n = lib().orbo_extract(self._h, _p(image), image.shape[0], image.shape[1], cap, _p(kps), _p(desc))
        assert n <= cap
        self._have = True
        return kps[:n].copy(), desc[:n].copy()

    def operator_kd(self, image):
        kps, desc = self.extract_arrays(image)
        tuples = [(float(k[0]), float(k[1]), float(k[2]), float(k[3]), float(k[4]), int(k[5])) for k in kps]
        return tuples, desc

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        lib().orbo_level_size(self._h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_bordered(self, l):
        w, h = self.level_size(l)
        out = np.empty((h + 38, w + 38), np.uint8)
        lib().orbo_level_bordered(self._h, l, _p(out))
        return out

    def level_blurred(self, l):
        w, h = self.level_size(l)
        out = np.zeros((h, w), np.uint8)
        lib().orbo_level_blurred(self._h, l, _p(out))
        return out

    def level_candidates(self, l):
        cap = 1 << 20
        out = np.empty((cap, 3), np.int32)
        n = lib().orbo_level_candidates(self._h, l, cap, _p(out))
        return out[:n].copy()

    def level_keypoints(self, l):
        cap = self.nfeatures + 1024
        out = np.empty((cap, 4), np.float32)
        n = lib().orbo_level_keypoints(self._h, l, cap, _p(out))
        return out[:n].copy()

    # getters, ORBextractor.h:62-86
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self._scale

    def GetScaleFactors(self):
        return [float(v) for v in self.sf]

    def GetInverseScaleFactors(self):
        return [float(v) for v in self.isf]

    def GetScaleSigmaSquares(self):
        return [float(v) for v in self.sig2]

    def GetInverseScaleSigmaSquares(self):
        return [float(v) for v in self.isig2]

    def GetImagePyramid(self):
        """The step-ignoring caster view (opencv_type_casters.h:205-240)."""
        out = []
        for l in range(self.nlevels):
            w, h = self.level_size(l)
            v = np.empty((h, w), np.uint8)
            lib().orbo_level_caster_view(self._h, l, _p(v))
            out.append(v)
        return out


# ---------------------------------------------------------------- stereo (C restatement)
def median_cull(uR, dep, sad):
    """Upstream ORB-SLAM2's outlier cull at the end of ComputeStereoMatches (NOT in the reference):
    sort the accepted matches by SAD minimum, median = element [n // 2], thDist = 1.5f * 1.4f * median, drop dist >= thDist."""
    uR, dep = uR.copy(), dep.copy()
    idx = np.nonzero(sad >= 0)[0]
    if len(idx):
        d = np.sort(sad[idx])
        th = np.float32(np.float32(1.5) * np.float32(1.4)) * np.float32(d[len(d) // 2])
        drop = idx[~(sad[idx].astype(np.float32) < th)]
        uR[drop] = -1
        dep[drop] = -1
    return uR, dep


def stereo(kpsL, descL, kpsR, descR, sf, isf, pyrL, pyrR, mbf, fx, with_sad=False):
    """kps*: [n,3] float32 (x, y, octave); pyr*: list of caster views (or of true level images for the "dense pyramid"
    option).  Returns uRight, depth, bestIdx, bestDist (+ the SAD minima with with_sad=True)."""
    kpsL = np.ascontiguousarray(kpsL, np.float32)
    kpsR = np.ascontiguousarray(kpsR, np.float32)
    descL = np.ascontiguousarray(descL, np.uint8)
    descR = np.ascontiguousarray(descR, np.uint8)
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
    bi = np.empty(n, np.int32)
    bd = np.empty(n, np.int32)
    sad = np.empty(n, np.int32)
    rc = lib().orbo_stereo_ex(n, _p(kpsL), _p(descL), len(kpsR), _p(kpsR), _p(descR), L, _p(sf), _p(isf),
                              C.cast(PL, C.c_void_p), C.cast(PR, C.c_void_p), _p(lw), _p(lh),
                              float(mbf), float(np.float32(fx)), _p(uR), _p(dep), _p(bi), _p(bd), _p(sad))
    if rc != 0:
        raise IndexError("stereo oracle: a keypoint row/window leaves the pyramid view (the reference raises here)")
    if with_sad:
        return uR, dep, bi, bd, sad
    return uR, dep, bi, bd
