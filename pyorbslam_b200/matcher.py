"""SURVEY.md 8(f) rank 1, first piece: `ORBMatcher.search_by_BoW_kf_f` / `search_by_BoW_kf_kf`
(reference ORBMatcher.py:21-213) with the Hamming distances taken from one GPU all-pairs matrix
(`b200orb_hamming_matrix`) instead of one Python `bin().count` call per candidate pair (ORBMatcher.py:12-14).

The greedy, order-dependent matching logic (skip already matched features, best / second-best ratio test, rotation
histogram with its np.argsort tie behaviour, python round()) is restated unchanged on the host -- only the distance
lookup differs -- so results are identical to the reference's.

    import pyorbslam_b200.matcher as m
    m.install_matcher(ORBMatcher)        # the reference's class; other methods stay untouched
"""
import ctypes as C

import numpy as np

from . import _lib

TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30      # ORBMatcher.py:3-5


def hamming_matrix(a, b, device=0):
    """uint16[nA, nB] of popcount(a[i] ^ b[j]) for uint8[n, 32] descriptor arrays."""
    a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
    b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
    out = np.empty((len(a), len(b)), np.uint16)
    l = _lib.lib()
    l.b200orb_hamming_matrix.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    if len(a) and len(b):
        _lib.check(l.b200orb_hamming_matrix(int(device), a.ctypes.data, len(a), b.ctypes.data, len(b), out.ctypes.data))
    return out


def _three_maxima(rot_hist):
    counts = [len(h) for h in rot_hist]
    return np.argsort(counts)[::-1][:3]          # compute_three_maxima, ORBMatcher.py:16-19 (same call, same tie order)


def _rot_bin(rot):
    if rot < 0.0:
        rot += 360.0
    b = round(rot * (1.0 / HISTO_LENGTH))
    return 0 if b == HISTO_LENGTH else b


def search_by_BoW_kf_f(self, kf, frame):
    """ORBMatcher.search_by_BoW_kf_f, ORBMatcher.py:21-118."""
    mps_kf = kf.get_map_point_matches()
    matches = [None] * frame.N
    D = hamming_matrix(kf.mDescriptors, frame.mDescriptors).tolist() if frame.N and len(mps_kf) else []
    n_matches = 0
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    fv_kf, fv_f = kf.mFeatVec, frame.mFeatVec
    it_kf, it_f = iter(fv_kf), iter(fv_f)
    try:
        node_kf, node_f = next(it_kf), next(it_f)
        while True:
            if node_kf == node_f:
                idx_f_list = fv_f[node_f]
                for i_kf in fv_kf[node_kf]:
                    mp = mps_kf[i_kf]
                    if not mp or mp.is_bad():
                        continue
                    row = D[i_kf]
                    best1, best_f, best2 = 256, -1, 256
                    for i_f in idx_f_list:
                        if matches[i_f]:
                            continue
                        d = row[i_f]
                        if d < best1:
                            best2, best1, best_f = best1, d, i_f
                        elif d < best2:
                            best2 = d
                    if best1 <= TH_LOW and float(best1) < self.mfNNratio * float(best2):
                        matches[best_f] = mp
                        if self.mbCheckOrientation:
                            b = _rot_bin(kf.mvKeysUn[i_kf].angle - frame.mvKeys[best_f].angle)
                            assert 0 <= b < HISTO_LENGTH
                            rot_hist[b].append(best_f)
                        n_matches += 1
                node_kf, node_f = next(it_kf), next(it_f)
            elif node_kf < node_f:
                node_kf = next(it_kf)
            else:
                node_f = next(it_f)
    except StopIteration:
        pass
    if self.mbCheckOrientation:
        keep = _three_maxima(rot_hist)
        for i in range(HISTO_LENGTH):
            if i in keep:
                continue
            for idx in rot_hist[i]:
                matches[idx] = None
                n_matches -= 1
    return n_matches, matches


def search_by_BoW_kf_kf(self, kf1, kf2):
    """ORBMatcher.search_by_BoW_kf_kf, ORBMatcher.py:120-213."""
    mps1, mps2 = kf1.get_map_point_matches(), kf2.get_map_point_matches()
    matches12 = [None] * len(mps1)
    matched2 = [False] * len(mps2)
    D = hamming_matrix(kf1.mDescriptors, kf2.mDescriptors).tolist() if len(mps1) and len(mps2) else []
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    n_matches = 0
    it1, it2 = iter(kf1.mFeatVec.items()), iter(kf2.mFeatVec.items())
    try:
        f1, f2 = next(it1), next(it2)
        while True:
            if f1[0] == f2[0]:
                for i1 in f1[1]:
                    mp1 = mps1[i1]
                    if not mp1 or mp1.is_bad():
                        continue
                    row = D[i1]
                    best1, best_i2, best2 = 256, -1, 256
                    for i2 in f2[1]:
                        mp2 = mps2[i2]
                        if matched2[i2] or not mp2 or mp2.is_bad():
                            continue
                        d = row[i2]
                        if d < best1:
                            best2, best1, best_i2 = best1, d, i2
                        elif d < best2:
                            best2 = d
                    if best1 < TH_LOW and best1 < self.mfNNratio * best2:
                        matches12[i1] = mps2[best_i2]
                        matched2[best_i2] = True
                        if self.mbCheckOrientation:
                            rot_hist[_rot_bin(kf1.mvKeysUn[i1].angle - kf2.mvKeysUn[best_i2].angle)].append(i1)
                        n_matches += 1
                f1, f2 = next(it1), next(it2)
            elif f1[0] < f2[0]:
                f1 = next(it1)
            else:
                f2 = next(it2)
    except StopIteration:
        pass
    if self.mbCheckOrientation:
        keep = _three_maxima(rot_hist)
        for i in range(HISTO_LENGTH):
            if i in keep:
                continue
            for idx in rot_hist[i]:
                matches12[idx] = None
                n_matches -= 1
    return n_matches, matches12


def install_matcher(matcher_cls):
    """Patch the two BoW searches of the reference's ORBMatcher class; returns the originals."""
    orig = (matcher_cls.search_by_BoW_kf_f, matcher_cls.search_by_BoW_kf_kf)
    matcher_cls.search_by_BoW_kf_f = search_by_BoW_kf_f
    matcher_cls.search_by_BoW_kf_kf = search_by_BoW_kf_kf
    return orig
