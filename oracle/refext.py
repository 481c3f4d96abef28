"""TEST INFRASTRUCTURE -- ctypes front for oracle/_ref/libref_{canonical,shipped}.so: the reference's own
ORBextractor.cpp compiled unmodified against oracle/cvshim (recipe: oracle/Makefile, target `ref`).

canonical = -O2 -ffp-contract=off + monotone list-node allocator (reproducible; the parity anchor)
shipped   = -O3 -march=x86-64-v3 -DNDEBUG + glibc malloc (what a user runs; the timed CPU baseline)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def available(kind="canonical"):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_{kind}.so"))


def _lib(kind):
    if kind not in _LIBS:
        l = C.CDLL(os.path.join(_HERE, "_ref", f"libref_{kind}.so"))
        l.ref_create.restype = C.c_void_p
        l.ref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        l.ref_destroy.argtypes = [C.c_void_p]
        l.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.ref_levels.argtypes = [C.c_void_p]
        l.ref_tables.argtypes = [C.c_void_p] * 5
        l.ref_level_size.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        l.ref_level_caster_view.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.ref_level_roi.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.ref_arena_overflows.restype = C.c_long
        _LIBS[kind] = l
    return _LIBS[kind]


class RefExtractor:
    """pyORBExtractor.ORBextractor surface (orb_extractor.cpp:22-38) over the compiled reference."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, kind="canonical"):
        self._l = _lib(kind)
        self._h = self._l.ref_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        self.nlevels = int(nlevels)
        self.nfeatures = int(nfeatures)
        self._scale = float(np.float32(scaleFactor))
        L = self.nlevels
        self.sf, self.isf, self.sig2, self.isig2 = (np.empty(L, np.float32) for _ in range(4))
        self._l.ref_tables(self._h, *(a.ctypes.data for a in (self.sf, self.isf, self.sig2, self.isig2)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.ref_destroy(self._h)
            self._h = None

    def extract_arrays(self, image):
        image = np.ascontiguousarray(image, np.uint8)
        cap = self.nfeatures + 66 * self.nlevels
        kps = np.empty((cap, 6), np.float32)
        desc = np.empty((cap, 32), np.uint8)
        n = self._l.ref_extract(self._h, image.ctypes.data, image.shape[0], image.shape[1], cap, kps.ctypes.data, desc.ctypes.data)
        assert n <= cap
        if self._l.ref_arena_overflows():
            raise MemoryError("reference list-node arena overflowed; tie order no longer canonical")
        return kps[:n].copy(), desc[:n].copy()

    def operator_kd(self, image):
        kps, desc = self.extract_arrays(image)
        return [(float(k[0]), float(k[1]), float(k[2]), float(k[3]), float(k[4]), int(k[5])) for k in kps], desc

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        self._l.ref_level_size(self._h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_roi(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        self._l.ref_level_roi(self._h, l, out.ctypes.data)
        return out

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
        out = []
        for l in range(self.nlevels):
            w, h = self.level_size(l)
            v = np.empty((h, w), np.uint8)
            self._l.ref_level_caster_view(self._h, l, v.ctypes.data)
            out.append(v)
        return out
