"""Multi-GPU front-end of one node: stereo pairs are independent, so a job is cut into contiguous frame ranges,
one per GPU (`sharding.shard_range`); every GPU has its own engine, pinned staging and host thread, and all of them write their
rows of ONE set of result arrays.  There is no data-path collective and nothing crosses NVLink."""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib
from .batch import StereoFrontend
from .sharding import shard_range


class StereoFrontendMulti:
    """`StereoFrontend.run_host` over several GPUs of one process.

    devices: CUDA device indices (default: all visible).  The ctypes call releases the GIL, so the per-GPU host threads really run
    side by side; each engine chunks, uploads, computes and downloads its own shard (b200orb_batch_run_host_shard)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, H, W, max_pairs, devices=None):
        if devices is None:
            devices = list(range(_lib.device_count()))
        if not devices:
            raise _lib.B200OrbError("no CUDA device: pyorbslam_b200 has no CPU fallback")
        self.devices = [int(d) for d in devices]
        self.engines = [StereoFrontend(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, H, W, max_pairs, device=d) for d in self.devices]
        self.H, self.W = int(H), int(W)
        self.capacity = self.engines[0].capacity
        self._pool = ThreadPoolExecutor(max_workers=len(self.devices), thread_name_prefix="b200orb-gpu")
        self.last_pair_status = None

    def alloc_outputs(self, n_pairs):
        return self.engines[0].alloc_outputs(n_pairs, pinned_host=True)

    def shards(self, n_pairs):
        """[(device, first_pair, stop_pair)] -- the frame range every GPU owns for a job of n_pairs pairs."""
        G = len(self.devices)
        return [(self.devices[r],) + shard_range(n_pairs, r, G) for r in range(G)]

    def run_host(self, left, right, mbf, fx, out=None):
        """left/right: uint8 HOST tensors or arrays [n, H, W] (pinned recommended).  Returns host tensors laid out exactly like
        StereoFrontend.run_host's (kps [2, n, C, 6], desc [2, n, C, 32], nkp [2, n], uRight / depth / matchIdx [n, C])."""
        lt = left if isinstance(left, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(left))
        rt = right if isinstance(right, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(right))
        if lt.is_cuda or rt.is_cuda or lt.dtype != torch.uint8 or rt.dtype != torch.uint8 or not lt.is_contiguous() or not rt.is_contiguous():
            raise ValueError("run_host takes contiguous uint8 host tensors")
        n = lt.shape[0]
        if tuple(lt.shape[1:]) != (self.H, self.W) or lt.shape != rt.shape or n < 1:
            raise ValueError(f"expected [n >= 1, {self.H}, {self.W}] for both views, got {tuple(lt.shape)} / {tuple(rt.shape)}")
        if out is None:
            out = self.alloc_outputs(n)
        else:
            self.engines[0]._check_out(out, n, on_device=False)
        flags = np.zeros(n, np.int32)
        HW = self.H * self.W
        lock = threading.Lock()
        errors = []

        def work(r):
            eng = self.engines[r]
            a, b = shard_range(n, r, len(self.engines))
            if b <= a:
                return
            rc = _lib.lib().b200orb_batch_run_host_shard(eng._h, lt.data_ptr() + a * HW, rt.data_ptr() + a * HW, b - a, float(mbf), float(np.float32(fx)),
                                                         out["kps"].data_ptr(), out["desc"].data_ptr(), out["nkp"].data_ptr(),
                                                         out["uRight"].data_ptr(), out["depth"].data_ptr(), out["matchIdx"].data_ptr(), n, a)
            if rc == _lib.E_RANGE:
                _lib.lib().b200orb_batch_status_host(eng._h, flags[a:b].ctypes.data, b - a)
            if rc != 0:
                with lock:        # last_error is per thread: fetch the message here
                    errors.append((rc, _lib.lib().b200orb_last_error().decode()))
        list(self._pool.map(work, range(len(self.engines))))
        self.last_out, self.last_pair_status = out, flags
        for rc, msg in errors:
            if rc != _lib.E_RANGE:
                raise _lib.B200OrbError(msg) if rc != _lib.E_ARG else ValueError(msg)
        if errors:
            raise IndexError(errors[0][1])
        return out

    def close(self):
        self._pool.shutdown(wait=True)
        self.engines = []
