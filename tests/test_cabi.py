"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200orb.h
declares, its host-side tables equal the oracle's, and compute calls fail loudly without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O
from pyorbslam_b200 import ORBextractor, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200orb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200orb_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_constructor_tables_match_oracle_and_reference_known_answers():
    e = ORBextractor(nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)   # keyword names of orb_extractor.cpp:23
    o = O.OracleExtractor(2000, 1.2, 8, 20, 7)
    assert e.GetLevels() == 8
    assert e.GetScaleFactor() == float(np.float32(1.2))
    assert e.GetScaleFactors() == o.GetScaleFactors()
    assert e.GetInverseScaleFactors() == o.GetInverseScaleFactors()
    assert e.GetScaleSigmaSquares() == o.GetScaleSigmaSquares()
    assert e.GetInverseScaleSigmaSquares() == o.GetInverseScaleSigmaSquares()
    assert e.features_per_level() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert all(isinstance(v, float) for v in e.GetScaleFactors())


def test_bad_parameters_raise():
    with pytest.raises(ValueError):
        ORBextractor(100, 1.2, 0, 20, 7)
    with pytest.raises(ValueError):
        ORBextractor(100, 1.0, 8, 20, 7)
    with pytest.raises(ValueError):
        ORBextractor(100, 1.2, 17, 20, 7)


def test_input_contract_errors_match_the_caster():
    e = ORBextractor(100, 1.2, 4, 20, 7)
    with pytest.raises(RuntimeError):      # opencv_type_casters.h:181-184
        e.operator_kd(np.zeros(10, np.uint8))
    with pytest.raises(RuntimeError):      # opencv_type_casters.h:195-197
        e.operator_kd(np.zeros((10, 10), np.float64))
    with pytest.raises(RuntimeError):      # CV_8UC1 only (ORBextractor.cpp:1049)
        e.operator_kd(np.zeros((10, 10, 3), np.uint8))
    with pytest.raises(_lib.B200OrbError):  # GetImagePyramid before any image
        e.GetImagePyramid()


def test_empty_image_returns_nothing_like_the_reference():
    e = ORBextractor(100, 1.2, 4, 20, 7)
    kps, desc = e.operator_kd(np.zeros((0, 0), np.uint8))   # ORBextractor.cpp:1045-1046
    assert kps == [] and desc.size == 0


@pytest.mark.skipif(_lib.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    e = ORBextractor(100, 1.2, 4, 20, 7)
    with pytest.raises(_lib.B200OrbError):
        e.operator_kd(np.zeros((100, 200), np.uint8))
    from pyorbslam_b200.stereo import stereo_host
    with pytest.raises(_lib.B200OrbError):
        stereo_host(np.zeros((1, 3), np.float32), np.zeros((1, 32), np.uint8), np.zeros((1, 3), np.float32), np.zeros((1, 32), np.uint8),
                    [1.0], [1.0], [np.zeros((50, 50), np.uint8)], [np.zeros((50, 50), np.uint8)], 100.0, 300.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyorbslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liborb_oracle" not in txt and "oracle/_ref" not in txt, f


def test_run_host_chunk_schedule_host_logic():
    """b200orb_host_chunk_schedule (pure host logic behind b200orb_batch_run_host): every pair lands in exactly one chunk, no chunk
    exceeds the engine, half-capacity chunks on two lanes, a C/4, C/2 ramp at both ends of jobs of at least four chunks."""
    rng = np.random.default_rng(0)
    for _ in range(300):
        P = int(rng.integers(1, 300)); n = int(rng.integers(1, 5000)); lanes = int(rng.integers(1, 3))
        s = _lib.chunk_schedule(P, lanes, n)
        C = P // 2 if (lanes == 2 and P >= 16) else P
        assert sum(s) == n and min(s) >= 1 and max(s) <= C
        if n >= 4 * C and C >= 8:
            assert s[:2] == [C // 4, C // 2] and s[-1] <= C // 2          # ramp up, taper down
            assert s[2:-3].count(C) >= len(s[2:-3]) - 2              # full chunks in between (the last ones may be the remainder / taper)
        else:
            assert s[:-1] == [C] * (len(s) - 1)
    # the bench's end-to-end job: 2048 pairs through a 128-pair engine
    assert _lib.chunk_schedule(128, 2, 2048) == [16, 32] + [64] * 30 + [32, 32, 16]
    assert _lib.chunk_schedule(128, 1, 2048) == [32, 64] + [128] * 14 + [64, 64, 32]
    assert _lib.chunk_schedule(2, 2, 3) == [2, 1]
    with pytest.raises(ValueError):
        _lib.chunk_schedule(0, 1, 5)
