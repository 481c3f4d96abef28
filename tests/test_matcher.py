"""SURVEY.md 8(f) rank 1, first piece -- ORBMatcher.search_by_BoW_kf_f / search_by_BoW_kf_kf (ORBMatcher.py:21-213).
Golden vectors come from the reference's own ORBMatcher class (tests/golden/make_golden.py::matcher_case)."""
import os
import types

import numpy as np
import pytest

from oracle import matcher_py as M

CASES = {"a": (0.7, True), "b": (1, True), "c": (0.9, False)}


def _uids(v):
    return np.array([-1 if p is None else p.uid for p in v], np.int64)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_restatement_vs_reference_orbmatcher(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "matcher_small.npz"))
    ratio, ori = CASES[tag]
    A, B = M.make_case()
    n, v = M.bow_kf_f(A, B, ratio, ori)
    assert n == int(g[f"kf_f_n_{tag}"]) and np.array_equal(_uids(v), g[f"kf_f_{tag}"])
    A, B = M.make_case()
    n, v = M.bow_kf_kf(A, B, ratio, ori)
    assert n == int(g[f"kf_kf_n_{tag}"]) and np.array_equal(_uids(v), g[f"kf_kf_{tag}"])


@pytest.mark.gpu
def test_hamming_matrix_vs_numpy():
    from pyorbslam_b200.matcher import hamming_matrix
    rng = np.random.default_rng(0)
    for na, nb in [(1, 1), (7, 300), (513, 129), (2004, 2007)]:
        a = rng.integers(0, 256, (na, 32), dtype=np.uint8)
        b = rng.integers(0, 256, (nb, 32), dtype=np.uint8)
        b[0] = ~a[0]                                            # distance 256 needs the 16-bit output
        ref = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2).sum(2).astype(np.uint16)
        assert np.array_equal(hamming_matrix(a, b), ref)
    assert hamming_matrix(np.zeros((0, 32), np.uint8), b).shape == (0, nb)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_gpu_matcher_vs_reference_orbmatcher(golden_dir, tag):
    from pyorbslam_b200.matcher import install_matcher
    g = np.load(os.path.join(golden_dir, "matcher_small.npz"))
    ratio, ori = CASES[tag]

    class Matcher:          # stands for the reference's ORBMatcher: same two attributes, methods patched in
        def __init__(self, nnratio, checkOri):
            self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri

        def search_by_BoW_kf_f(self, kf, f):
            raise AssertionError("unpatched")

        def search_by_BoW_kf_kf(self, a, b):
            raise AssertionError("unpatched")
    install_matcher(Matcher)
    m = Matcher(ratio, ori)
    A, B = M.make_case()
    n, v = m.search_by_BoW_kf_f(A, B)
    assert n == int(g[f"kf_f_n_{tag}"]) and np.array_equal(_uids(v), g[f"kf_f_{tag}"])
    A, B = M.make_case()
    n, v = m.search_by_BoW_kf_kf(A, B)
    assert n == int(g[f"kf_kf_n_{tag}"]) and np.array_equal(_uids(v), g[f"kf_kf_{tag}"])
    # a larger seeded case against the oracle restatement
    A, B = M.make_case(seed=9, n=900, n_nodes=20)
    n1, v1 = m.search_by_BoW_kf_f(A, B)
    A2, B2 = M.make_case(seed=9, n=900, n_nodes=20)
    n2, v2 = M.bow_kf_f(A2, B2, ratio, ori)
    assert n1 == n2 and np.array_equal(_uids(v1), _uids(v2))


# ---------------------------------------------------------------- projection searches (ORBMatcher.py:215-393)
PCASES = {"p": (0.9, True, 15, (0.05, 0.0, 0.3)), "q": (0.8, False, 7, (0.0, 0.02, -0.9)), "r": (1, True, 10, (0.4, 0.0, 0.0))}


@pytest.mark.parametrize("tag", sorted(PCASES))
def test_oracle_projection_vs_reference_orbmatcher(golden_dir, tag):
    """The restated searches AND the restated Frame.get_features_in_area / assign_features_to_grid against the
    reference's own classes (golden)."""
    g = np.load(os.path.join(golden_dir, "matcher_small.npz"))
    ratio, ori, th, motion = PCASES[tag]
    cur, last, local = M.make_projection_case(motion=motion)
    n = M.projection_f_f(cur, last, th, ori)
    assert n == int(g[f"f_f_n_{tag}"]) and np.array_equal(_uids(cur.mvpMapPoints), g[f"f_f_{tag}"])
    cur, last, local = M.make_projection_case(motion=motion)
    n = M.projection_f_p(cur, local, float(th) / 5, ratio)
    assert n == int(g[f"f_p_n_{tag}"]) and np.array_equal(_uids(cur.mvpMapPoints), g[f"f_p_{tag}"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(PCASES))
def test_gpu_projection_vs_reference_orbmatcher(golden_dir, tag):
    from pyorbslam_b200.matcher import install_matcher
    g = np.load(os.path.join(golden_dir, "matcher_small.npz"))
    ratio, ori, th, motion = PCASES[tag]

    class Matcher:
        def __init__(self, nnratio, checkOri):
            self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri

        def radius_by_viewing_cos(self, c):             # ORBMatcher.py:285-289, unpatched in the reference too
            return 2.5 if c > 0.998 else 4.0
    assert install_matcher(Matcher) == {}                # this stand-in had none of the four; the reference's class returns its originals
    assert all(hasattr(Matcher, n) for n in ("search_by_BoW_kf_f", "search_by_BoW_kf_kf", "search_by_projection_f_f", "search_by_projection_f_p"))
    m = Matcher(ratio, ori)
    cur, last, local = M.make_projection_case(motion=motion)
    n = m.search_by_projection_f_f(cur, last, th)
    assert n == int(g[f"f_f_n_{tag}"]) and np.array_equal(_uids(cur.mvpMapPoints), g[f"f_f_{tag}"])
    cur, last, local = M.make_projection_case(motion=motion)
    n = m.search_by_projection_f_p(cur, local, float(th) / 5)
    assert n == int(g[f"f_p_n_{tag}"]) and np.array_equal(_uids(cur.mvpMapPoints), g[f"f_p_{tag}"])
    # a larger seeded scene against the oracle restatement
    cur, last, local = M.make_projection_case(seed=11, n=1800, motion=motion)
    n1 = m.search_by_projection_f_f(cur, last, th)
    c2, l2, _ = M.make_projection_case(seed=11, n=1800, motion=motion)
    assert n1 == M.projection_f_f(c2, l2, th, ori) and np.array_equal(_uids(cur.mvpMapPoints), _uids(c2.mvpMapPoints))
    # degenerate inputs: no map points at all / nothing in view
    cur, last, local = M.make_projection_case(seed=3, n=50, motion=motion)
    last.mvpMapPoints = [None] * last.N
    assert m.search_by_projection_f_f(cur, last, th) == 0
    for p in local:
        p.mbTrackInView = False
    assert m.search_by_projection_f_p(cur, local, 1.0) == 0


# ---------------------------------------------------------------- the pieces behind the projection searches
def _greedy_py(start, idx, dist, ok, occ, marks, koct=None, th=100, ratio=None):
    """Pure-Python statement of ORBMatcher.py:236-281 (ratio given) / :335-366 (ratio None) on candidate lists."""
    occ = occ.copy()
    best = np.full(len(start) - 1, -1, np.int32)
    for q in range(len(start) - 1):
        b1, l1, b2, l2, bi = 256, -1, 256, -1, -1
        for c in range(start[q], start[q + 1]):
            j = idx[c]
            if occ[j] or not ok[c]:
                continue
            d = dist[c]
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, (koct[j] if koct is not None else 0), j
            elif d < b2:
                l2, b2 = (koct[j] if koct is not None else 0), d
        if b1 <= th and bi >= 0:
            if ratio is not None and l1 == l2 and b1 > ratio * b2:
                continue
            best[q] = bi
            occ[bi] = marks[q]
    return best, occ


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_greedy_selection_host_functions_vs_python(seed):
    """b200orb_greedy_project_ff / _fp are host code behind the C ABI: checked here without a GPU."""
    import ctypes as C
    from pyorbslam_b200 import _lib
    rng = np.random.default_rng(seed)
    N, Mq = 300, 400
    cnt = rng.integers(0, 12, Mq)
    start = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    idx = rng.integers(0, N, start[-1]).astype(np.int32)
    dist = rng.choice([5, 30, 60, 99, 100, 101, 140], start[-1]).astype(np.int32)       # many ties, values around TH_HIGH
    ok = (rng.random(start[-1]) < 0.85).astype(np.uint8)
    occ0 = (rng.random(N) < 0.2).astype(np.uint8)
    marks = (rng.random(Mq) < 0.6).astype(np.uint8)
    koct = rng.integers(0, 3, N).astype(np.int32)
    l = _lib.lib()
    l.b200orb_greedy_project_ff.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    l.b200orb_greedy_project_fp.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                                                            C.c_void_p]
    occ, best = occ0.copy(), np.empty(Mq, np.int32)
    _lib.check(l.b200orb_greedy_project_ff(Mq, start.ctypes.data, idx.ctypes.data, dist.ctypes.data, ok.ctypes.data, N, occ.ctypes.data,
                                           marks.ctypes.data, 100, best.ctypes.data))
    eb, eo = _greedy_py(start, idx, dist, ok, occ0, marks)
    assert np.array_equal(best, eb) and np.array_equal(occ, eo)
    for ratio in (0.6, 0.9, 1.0):
        occ, best = occ0.copy(), np.empty(Mq, np.int32)
        _lib.check(l.b200orb_greedy_project_fp(Mq, start.ctypes.data, idx.ctypes.data, dist.ctypes.data, ok.ctypes.data, N, occ.ctypes.data,
                                               marks.ctypes.data, koct.ctypes.data, 100, ratio, best.ctypes.data))
        eb, eo = _greedy_py(start, idx, dist, ok, occ0, marks, koct, 100, ratio)
        assert np.array_equal(best, eb) and np.array_equal(occ, eo)
    bad = idx.copy(); bad[0] = N                                  # a candidate that is not a feature index
    if start[-1] > 0 and cnt[0] > 0:
        with pytest.raises(IndexError):
            _lib.check(l.b200orb_greedy_project_ff(Mq, start.ctypes.data, bad.ctypes.data, dist.ctypes.data, ok.ctypes.data, N,
                                                   occ0.copy().ctypes.data, marks.ctypes.data, 100, best.ctypes.data))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_area_query_and_distances_vs_get_features_in_area(dtype):
    """b200orb_area_hamming against the restated Frame.get_features_in_area (itself pinned to the reference's by the golden
    vectors above), query by query, in both arithmetic modes; includes windows that leave the grid and empty ones."""
    from pyorbslam_b200.matcher import _area_hamming
    cur, last, local = M.make_projection_case(seed=21, n=900)
    rng = np.random.default_rng(5)
    Mq = 700
    qx = rng.uniform(-80, 1320, Mq).astype(dtype)
    qy = rng.uniform(-60, 440, Mq).astype(dtype)
    qr = rng.choice([0.0, 3.5, 10.0, 24.9, 64.5, 300.0], Mq)
    lv = rng.integers(-1, 8, (Mq, 2)).astype(np.int32)
    lv[::7] = (0, -1)                                             # "no level check" combination
    qd = rng.integers(0, 256, (Mq, 32), dtype=np.uint8)
    start, idx, dist = _area_hamming(cur, qx, qy, qr, lv, qd)
    assert start[0] == 0 and start[-1] == len(idx) == len(dist)
    for q in range(Mq):
        x, y = (qx[q], qy[q]) if dtype is np.float32 else (float(qx[q]), float(qy[q]))      # np.float32 scalar / Python float
        ref = M.features_in_area(cur, x, y, float(qr[q]), int(lv[q, 0]), int(lv[q, 1]))
        got = idx[start[q]:start[q + 1]].tolist()
        assert got == ref, (q, got[:5], ref[:5])
        if ref:
            d = np.unpackbits(cur.mDescriptors[ref] ^ qd[q], axis=1).sum(1)
            assert np.array_equal(dist[start[q]:start[q + 1]], d)
    assert len(idx) > 1000


@pytest.mark.gpu
def test_gpu_area_query_error_paths():
    """Too-small output buffers report the size needed; cell ranges outside the grid and broken grids are refused."""
    import ctypes as C
    from pyorbslam_b200 import _lib
    from pyorbslam_b200.matcher import _grid_csr
    cur, _, _ = M.make_projection_case(seed=3, n=200)
    kxy = np.array([k.pt for k in cur.mvKeysUn], np.float32)
    koct = np.array([k.octave for k in cur.mvKeysUn], np.int32)
    kd = np.ascontiguousarray(cur.mDescriptors)
    cs, ci = _grid_csr(cur)
    Mq = 4
    qxyr = np.array([[600.0, 180.0, 5000.0]] * Mq)
    qlvl = np.array([[0, -1]] * Mq, np.int32)
    qcell = np.array([[0, 63, 0, 47]] * Mq, np.int32)
    qd = np.zeros((Mq, 32), np.uint8)
    start = np.zeros(Mq + 1, np.int32)
    total = C.c_int32(0)
    l = _lib.lib()
    l.b200orb_area_hamming.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 2 + \
                                      [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p]

    def call(cell_start=cs, cells=qcell, cap=0, idx=None, dist=None):
        return l.b200orb_area_hamming(0, 0, 200, kxy.ctypes.data, koct.ctypes.data, kd.ctypes.data, 64, 48, cell_start.ctypes.data,
                                      ci.ctypes.data, Mq, qxyr.ctypes.data, qlvl.ctypes.data, cells.ctypes.data, qd.ctypes.data,
                                      start.ctypes.data, idx, dist, cap, C.byref(total))
    assert call() != 0 and total.value > 0                      # every feature of the frame, four times: does not fit in cap = 0
    need = total.value
    idx, dist = np.empty(need, np.int32), np.empty(need, np.int32)
    assert call(cap=need, idx=idx.ctypes.data, dist=dist.ctypes.data) == 0 and start[-1] == need
    assert sorted(idx[:start[1]].tolist()) == sorted(ci.tolist())
    bad = qcell.copy(); bad[1] = (0, 64, 0, 47)
    with pytest.raises(IndexError):
        _lib.check(call(cells=bad))
    broken = cs.copy(); broken[5] = broken[4] - 1
    with pytest.raises(ValueError):
        _lib.check(call(cell_start=broken))
