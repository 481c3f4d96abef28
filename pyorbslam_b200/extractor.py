"""`ORBextractor`: the reference's pybind class (pyORBExtractor/orb_extractor.cpp:22-38) on the B200.

Same constructor keywords, same nine methods, same return types: `operator_kd(image)` returns
(list of (x, y, size, angle, response, octave) tuples, np.uint8[N,32]) -- owning copies, like the reference's
casters (opencv_type_casters.h:106-108, 205-240).  The results and the image pyramid also stay resident on the
device, so the patched `Frame.compute_stereo_matches` (stereo.py) can match without re-uploading anything."""
import ctypes as C
import weakref

import numpy as np

from . import _lib


class ORBextractor:
    _live = weakref.WeakSet()      # extractors whose last results may still be resident (bow.py looks descriptors up here)

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0, reuse_identical_input=True):
        """reuse_identical_input: the reference's Frame.copy() runs the whole Frame constructor again on the same two
        images (Frame.py:75-77, called at Tracking.py:267,306), i.e. it extracts every tracked frame twice.  When the image
        passed to operator_kd is byte-identical to the previous call's, the (deterministic) results still resident on the
        device are returned again instead of being recomputed; a byte compare of 0.5 MB costs ~25 us."""
        self._h = None
        self._reuse = bool(reuse_identical_input)
        self._last_image = None
        self._last_kps = None
        self._cached_desc = None
        self.reused_calls = 0
        h = C.c_void_p()
        _lib.check(_lib.lib().b200orb_extractor_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST),
                                                       int(minThFAST), int(device), C.byref(h)))
        self._h = h
        ORBextractor._live.add(self)
        self._nlevels = int(nlevels)
        self._last_desc = None      # identity token: the descriptor array handed out by the last operator_kd
        self._last_n = -1

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib().b200orb_extractor_destroy(self._h)
            except Exception:
                pass
            self._h = None

    # ---- getters, ORBextractor.h:62-86 ----
    def GetLevels(self):
        return int(_lib.lib().b200orb_get_levels(self._h))

    def GetScaleFactor(self):
        return float(_lib.lib().b200orb_get_scale_factor(self._h))

    def _tab(self, fn, dtype=np.float32):
        out = np.empty(self._nlevels, dtype)
        _lib.check(getattr(_lib.lib(), fn)(self._h, out.ctypes.data))
        return out

    def GetScaleFactors(self):
        return self._tab("b200orb_get_scale_factors").tolist()

    def GetInverseScaleFactors(self):
        return self._tab("b200orb_get_inverse_scale_factors").tolist()

    def GetScaleSigmaSquares(self):
        return self._tab("b200orb_get_scale_sigma_squares").tolist()

    def GetInverseScaleSigmaSquares(self):
        return self._tab("b200orb_get_inverse_scale_sigma_squares").tolist()

    def features_per_level(self):
        return self._tab("b200orb_get_features_per_level", np.int32).tolist()

    # ---- operator_kd, orb_extractor.cpp:31-38 ----
    def extract_arrays(self, image):
        """(kps float32[N,6], desc uint8[N,32]) -- the array form of operator_kd."""
        image = np.asarray(image)
        if image.ndim not in (2, 3):
            raise RuntimeError(f"Unsupported dim {image.ndim}, only support 2d, or 3-d")   # opencv_type_casters.h:181-184
        if image.dtype not in (np.uint8, np.int32, np.float32):
            raise RuntimeError("Unsupported type, only support uchar, int32, float")        # opencv_type_casters.h:195-197
        if image.ndim != 2 or image.dtype != np.uint8:
            # the reference asserts CV_8UC1 (ORBextractor.cpp:1049) but compiles the assert out (-DNDEBUG) and then
            # reads the buffer as if it were 8-bit gray; we refuse instead of reproducing undefined behaviour
            raise RuntimeError("image must be 8-bit single channel (CV_8UC1)")
        image = np.ascontiguousarray(image)
        if (self._reuse and self._last_image is not None and self._last_image.shape == image.shape
                and np.array_equal(self._last_image, image)):
            self.reused_calls += 1
            desc = self._cached_desc.copy()        # owning copies, like every call (the private cache is never handed out)
            self._last_desc = desc                 # identity token for the device-resident stereo path
            return self._last_kps.copy(), desc
        n = C.c_int(0)
        H, W = image.shape
        _lib.check(_lib.lib().b200orb_extract(self._h, image.ctypes.data, H, W, C.byref(n)))
        n = n.value
        kps = np.empty((n, 6), np.float32)
        desc = np.empty((n, 32), np.uint8) if n else np.zeros((0, 0), np.uint8)
        if n:
            _lib.check(_lib.lib().b200orb_get_results(self._h, kps.ctypes.data, desc.ctypes.data))
        self._last_desc = desc
        self._last_n = n
        if self._reuse:
            self._last_image = image.copy()
            self._last_kps = kps.copy()
            self._cached_desc = desc.copy()
        return kps, desc

    def operator_kd(self, image):
        kps, desc = self.extract_arrays(image)
        # list of (float, float, float, float, float, int) tuples like the KeyPoint caster (opencv_type_casters.h:107): the
        # columns become Python lists at C speed and zip() builds the tuples (~35 % faster than a structured array's tolist();
        # the 14 k Python objects of 2 000 tuples are what is left of the cost)
        t = kps.T
        return list(zip(t[0].tolist(), t[1].tolist(), t[2].tolist(), t[3].tolist(), t[4].tolist(), t[5].astype(np.int32).tolist())), desc

    # ---- GetImagePyramid, ORBextractor.h:84-86 through the Mat caster ----
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _lib.check(_lib.lib().b200orb_level_size(self._h, int(level), C.byref(w), C.byref(h)))
        return w.value, h.value

    def GetImagePyramid(self):
        sizes = [self.level_size(l) for l in range(self._nlevels)]
        buf = np.empty(sum(w * h for w, h in sizes), np.uint8)        # one owning buffer; the levels are views of it
        _lib.check(_lib.lib().b200orb_get_pyramid_all(self._h, buf.ctypes.data, buf.size))
        out, o = [], 0
        for w, h in sizes:
            out.append(buf[o:o + w * h].reshape(h, w))
            o += w * h
        return out

    # ---- diagnostics used by the parity tests ----
    def level_image(self, level, blurred=False):
        w, h = self.level_size(level)
        v = np.empty((h, w), np.uint8)
        _lib.check(_lib.lib().b200orb_get_level_image(self._h, int(level), int(bool(blurred)), v.ctypes.data))
        return v

    def level_candidates(self, level):
        n = C.c_int(0)
        _lib.check(_lib.lib().b200orb_get_level_candidates(self._h, int(level), 0, None, C.byref(n)))
        out = np.empty((max(n.value, 1), 3), np.int32)
        _lib.check(_lib.lib().b200orb_get_level_candidates(self._h, int(level), n.value, out.ctypes.data, C.byref(n)))
        return out[:n.value]
