"""Drop-in for the reference's extension module `pyORBExtractor` (pyORBExtractor/orb_extractor.cpp).

The reference does `sys.path.append("./pyORBExtractor/lib/"); from pyORBExtractor import ORBextractor`
(Tracking.py:18-19, pyORBExtractor/test.py:8-9).  Put THIS directory on sys.path (or copy/symlink this file to
<reference>/pyORBExtractor/lib/) and the same import yields the B200 extractor; a real module beats the
namespace-package candidate `<reference>/pyORBExtractor/`."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from pyorbslam_b200.extractor import ORBextractor  # noqa: E402,F401
