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


def _reference_cells(w, h):
    """The cell loops of ComputeKeyPointsOctTree (ORBextractor.cpp:770-806) restated: (iniX, iniY, cw, ch) per (row, column) with
    cw = ch = 0 for the cells the loops `continue` over; detection area = FAST window minus its 3-px rim."""
    f32 = np.float32
    min_b, max_bx, max_by = 16, w - 19 + 3, h - 19 + 3                      # EDGE_THRESHOLD - 3, cols/rows - EDGE_THRESHOLD + 3
    width, height = f32(max_bx - min_b), f32(max_by - min_b)
    n_cols, n_rows = int(width / f32(30)), int(height / f32(30))
    if n_cols <= 0 or n_rows <= 0:
        return []
    w_cell, h_cell = int(np.ceil(width / f32(n_cols))), int(np.ceil(height / f32(n_rows)))
    cells = []
    for i in range(n_rows):
        ini_y = min_b + i * h_cell
        max_y = min(ini_y + h_cell + 6, max_by)
        for j in range(n_cols):
            ini_x = min_b + j * w_cell
            max_x = min(ini_x + w_cell + 6, max_bx)
            skipped = ini_y >= max_by - 3 or ini_x >= max_bx - 6 or max_x - ini_x < 7 or max_y - ini_y < 7      # :793, :801, FAST needs 7 px
            cells.append((ini_x, ini_y, 0, 0) if skipped else (ini_x, ini_y, max_x - ini_x - 6, max_y - ini_y - 6))
    return cells


@pytest.mark.parametrize("H,W,params", [(376, 1241, (2000, 1.2, 8, 20, 7)), (200, 480, (800, 1.2, 5, 20, 7)), (131, 257, (300, 1.5, 4, 20, 7)),
                                        (90, 300, (500, 2.0, 3, 20, 7))])
def test_planned_fast_cells_are_the_reference_loops_cells(H, W, params):
    """b200orb_plan_cells (host logic: the table k_fast_cells walks) against the restated loops on the oracle's level sizes, and
    against what the reference itself found: every FAST candidate of the compiled reference lies in the detection area of exactly
    one planned cell, and no cell holds more candidates than its segment has room for."""
    from pyorbslam_b200.synthetic import make_kitti_like_pair
    img = make_kitti_like_pair(3, H, W)[0]
    o = O.OracleExtractor(*params)
    o.extract_arrays(img)
    cells = _lib.plan_cells(*params, H, W)
    k = 0
    for l in range(params[2]):
        w, h = o.level_size(l)
        ref = _reference_cells(w, h)
        mine = cells[k:k + len(ref)]
        assert len(mine) == len(ref) and (mine[:, 0] == l).all()
        for (x, y, cw, ch), m in zip(ref, mine):
            assert (int(m[3]), int(m[4])) == (cw, ch)
            if cw:
                assert (int(m[1]), int(m[2])) == (x, y)
        cand = o.level_candidates(l)                                        # x, y relative to (16, 16), score
        hits = np.zeros(len(cand), np.int32)
        caps = np.append(cells[k + 1:k + len(ref), 5], 0) - mine[:, 5] if len(ref) else []
        for n_c, m in enumerate(mine):
            if m[3] == 0:
                continue
            x0, y0 = m[1] + 3 - 16, m[2] + 3 - 16
            inside = (cand[:, 0] >= x0) & (cand[:, 0] < x0 + m[3]) & (cand[:, 1] >= y0) & (cand[:, 1] < y0 + m[4])
            hits += inside
            if n_c + 1 < len(mine):
                assert inside.sum() <= caps[n_c]
        assert (hits == 1).all()
        k += len(ref)
    assert k == len(cells)
