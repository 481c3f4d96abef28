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
