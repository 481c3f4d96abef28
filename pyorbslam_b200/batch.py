"""Batched stereo front-end (throughput API): N independent stereo pairs per call, everything on the device.

PyTorch is used only for device/pinned buffers and the stream handle; the work is libb200orb's kernels."""
import ctypes as C

import numpy as np
import torch

from . import _lib


class StereoFrontend:
    """What Tracking.grab_image_stereo -> Frame.__init__ computes per pair (Tracking.py:95-112, Frame.py:48-65):
    ORB extraction of left and right + compute_stereo_matches, for `max_pairs` pairs per launch sequence."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, H, W, max_pairs, device=0):
        self._h = None
        self.device = int(device)
        self.H, self.W = int(H), int(W)
        h = C.c_void_p()
        _lib.check(_lib.lib().b200orb_batch_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST),
                                                   self.H, self.W, int(max_pairs), self.device, C.byref(h)))
        self._h = h
        self.max_pairs = int(_lib.lib().b200orb_batch_max_pairs(h))
        self.capacity = int(_lib.lib().b200orb_batch_kp_capacity(h))

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib().b200orb_batch_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def set_stereo_options(self, median_cull=False, dense_pyramid=False):
        """Opt-in upstream-ORB-SLAM2 behaviour (not the reference's): see B200ORB_STEREO_* in include/b200orb.h."""
        _lib.check(_lib.lib().b200orb_batch_set_stereo_flags(self._h, (1 if median_cull else 0) | (2 if dense_pyramid else 0)))

    def set_copy_only(self, on=True):
        """Measurement aid: run_host moves its bytes but launches no kernel (see b200orb_batch_set_copy_only)."""
        _lib.check(_lib.lib().b200orb_batch_set_copy_only(self._h, int(bool(on))))

    def workspace_bytes(self):
        return int(_lib.lib().b200orb_batch_workspace_bytes(self._h))

    STAGES = ("border0", "resize_chain", "blur", "fast_cells", "octree", "orient_describe", "stereo")

    def profile(self, enable=True, max_calls=1024):
        """Record CUDA events between the kernels of the next run() calls (on the launching stream)."""
        _lib.check(_lib.lib().b200orb_batch_profile(self._h, int(bool(enable)), int(max_calls)))

    def profile_read(self):
        """-> ({stage: total ms}, calls, pairs) over the calls recorded since profile(); clears the record."""
        ms = (C.c_float * 7)()
        n = C.c_int(0)
        pairs = C.c_longlong(0)
        _lib.check(_lib.lib().b200orb_batch_profile_read(self._h, ms, C.byref(n), C.byref(pairs)))
        return {s: float(ms[i]) for i, s in enumerate(self.STAGES)}, n.value, pairs.value

    def stage_launches(self):
        """{stage: kernel launches per run() call}"""
        v = (C.c_int * 7)()
        _lib.check(_lib.lib().b200orb_batch_stage_launches(self._h, v))
        return {s: int(v[i]) for i, s in enumerate(self.STAGES)}

    def candidate_count(self, n_images):
        t = C.c_longlong(0)
        _lib.check(_lib.lib().b200orb_batch_candidate_count(self._h, int(n_images), C.byref(t)))
        return int(t.value)

    def alloc_outputs(self, n_pairs, pinned_host=False):
        C_ = self.capacity
        kw = dict(device="cpu", pin_memory=True) if pinned_host else dict(device=f"cuda:{self.device}")
        return {
            "kps": torch.empty((2, n_pairs, C_, 6), dtype=torch.float32, **kw),
            "desc": torch.empty((2, n_pairs, C_, 32), dtype=torch.uint8, **kw),
            "nkp": torch.empty((2, n_pairs), dtype=torch.int32, **kw),
            "uRight": torch.empty((n_pairs, C_), dtype=torch.float32, **kw),
            "depth": torch.empty((n_pairs, C_), dtype=torch.float32, **kw),
            "matchIdx": torch.empty((n_pairs, C_), dtype=torch.int32, **kw),
        }

    _OUT_SPEC = (("kps", torch.float32, lambda n, c: (2, n, c, 6)), ("desc", torch.uint8, lambda n, c: (2, n, c, 32)),
                 ("nkp", torch.int32, lambda n, c: (2, n)), ("uRight", torch.float32, lambda n, c: (n, c)),
                 ("depth", torch.float32, lambda n, c: (n, c)), ("matchIdx", torch.int32, lambda n, c: (n, c)))

    def _check_out(self, out, n, on_device):
        """A caller-supplied `out` is written through raw pointers: refuse anything that is not exactly what alloc_outputs(n) makes."""
        for name, dtype, shape in self._OUT_SPEC:
            t = out.get(name) if isinstance(out, dict) else None
            if not isinstance(t, torch.Tensor):
                raise ValueError(f"out[{name!r}] is missing")
            if t.dtype != dtype or tuple(t.shape) != shape(n, self.capacity) or not t.is_contiguous():
                raise ValueError(f"out[{name!r}] must be a contiguous {dtype} tensor of shape {shape(n, self.capacity)}, got {t.dtype} {tuple(t.shape)}")
            if on_device and not (t.is_cuda and t.device.index == self.device):
                raise ValueError(f"out[{name!r}] must live on cuda:{self.device}, got {t.device}")
            if not on_device and t.is_cuda:
                raise ValueError(f"out[{name!r}] must be a host tensor")

    def check_status(self, n_pairs, stream=None):
        """Synchronises `stream` (default: the current one) and raises IndexError if a pair of the last run() hit what the reference
        raises for (a row band / SAD window leaving the pyramid view, Frame.py:192,230-250).  Returns the per-pair flags."""
        st = (stream or torch.cuda.current_stream(torch.device("cuda", self.device))).cuda_stream
        flags = np.zeros(int(n_pairs), np.int32)
        self.last_pair_status = flags
        _lib.check(_lib.lib().b200orb_batch_status_device(self._h, int(n_pairs), C.c_void_p(st), flags.ctypes.data))
        return flags

    def run(self, left, right, mbf, fx, out=None):
        """left/right: uint8 CUDA tensors [n, H, W] (n <= max_pairs).  Asynchronous on the current stream; range errors
        (see check_status) are reported by check_status(n), not here."""
        if not (left.is_cuda and right.is_cuda) or left.dtype != torch.uint8 or right.dtype != torch.uint8:
            raise ValueError("run takes uint8 CUDA tensors")
        if not (left.is_contiguous() and right.is_contiguous()) or left.shape != right.shape:
            raise ValueError("left/right must be contiguous and of equal shape")
        n = left.shape[0]
        if tuple(left.shape[1:]) != (self.H, self.W) or n < 1 or n > self.max_pairs:
            raise ValueError(f"expected [n <= {self.max_pairs}, {self.H}, {self.W}], got {tuple(left.shape)}")
        if left.device.index != self.device or right.device.index != self.device:
            raise ValueError(f"inputs must live on cuda:{self.device}")
        if out is None:
            out = self.alloc_outputs(n)
        else:
            self._check_out(out, n, on_device=True)
        st = torch.cuda.current_stream(left.device).cuda_stream
        _lib.check(_lib.lib().b200orb_batch_run_device(self._h, left.data_ptr(), right.data_ptr(), n, float(mbf), float(np.float32(fx)),
                                                       out["kps"].data_ptr(), out["desc"].data_ptr(), out["nkp"].data_ptr(),
                                                       out["uRight"].data_ptr(), out["depth"].data_ptr(), out["matchIdx"].data_ptr(),
                                                       C.c_void_p(st)))
        return out

    def run_host(self, left, right, mbf, fx, out=None):
        """left/right: uint8 HOST tensors or arrays [n, H, W] (pinned recommended, n may exceed max_pairs).
        Uploads, processes and downloads chunk by chunk with copy/compute overlap; returns host tensors."""
        lt = left if isinstance(left, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(left))
        rt = right if isinstance(right, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(right))
        if lt.is_cuda or rt.is_cuda:
            raise ValueError("run_host takes host tensors; use run() for device tensors")
        if lt.dtype != torch.uint8 or rt.dtype != torch.uint8 or not lt.is_contiguous() or not rt.is_contiguous():
            raise ValueError(f"run_host needs contiguous uint8 tensors, got {lt.dtype}/{rt.dtype} contiguous={lt.is_contiguous()}/{rt.is_contiguous()}")
        n = lt.shape[0]
        if tuple(lt.shape[1:]) != (self.H, self.W) or lt.shape != rt.shape:
            raise ValueError(f"expected [n, {self.H}, {self.W}] for both views, got {tuple(lt.shape)} / {tuple(rt.shape)}")
        if out is None:
            out = self.alloc_outputs(n, pinned_host=True)
        else:
            self._check_out(out, n, on_device=False)
        rc = _lib.lib().b200orb_batch_run_host(self._h, lt.data_ptr(), rt.data_ptr(), n, float(mbf), float(np.float32(fx)),
                                               out["kps"].data_ptr(), out["desc"].data_ptr(), out["nkp"].data_ptr(),
                                               out["uRight"].data_ptr(), out["depth"].data_ptr(), out["matchIdx"].data_ptr())
        self.last_out = out                      # on a range error (IndexError) the other pairs' results are still valid
        if rc == _lib.E_RANGE:
            flags = np.zeros(n, np.int32)
            _lib.lib().b200orb_batch_status_host(self._h, flags.ctypes.data, n)
            self.last_pair_status = flags
        _lib.check(rc)
        self.last_pair_status = np.zeros(n, np.int32)
        return out
