"""pyorbslam_b200 -- B200-native stereo ORB front-end for pyOrbSLAM2 (hot path only).

  ORBextractor      drop-in for pyORBExtractor.ORBextractor            (extractor.py)
  install(Frame)    GPU Frame.compute_stereo_matches                    (stereo.py)
  StereoFrontend    batched, device-resident throughput API             (batch.py, needs torch)
  StereoFrontendMulti  the same over all GPUs of one process, frames sharded per GPU (multi.py)

All compute lives in libb200orb.so (hand-written sm_100a CUDA behind the C ABI in include/b200orb.h)."""
from .extractor import ORBextractor  # noqa: F401
from .stereo import compute_stereo_matches, install, stereo_host  # noqa: F401

__all__ = ["ORBextractor", "install", "compute_stereo_matches", "stereo_host", "StereoFrontend", "StereoFrontendMulti"]


def __getattr__(name):
    if name == "StereoFrontend":     # torch is imported lazily: the extractor object does not need it
        from .batch import StereoFrontend
        return StereoFrontend
    if name == "StereoFrontendMulti":
        from .multi import StereoFrontendMulti
        return StereoFrontendMulti
    raise AttributeError(name)
