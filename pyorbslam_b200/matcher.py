"""the tracking-side searches of the reference's ORBMatcher -- `search_by_BoW_kf_f`,
`search_by_BoW_kf_kf` (ORBMatcher.py:21-213), `search_by_projection_f_p` (ORBMatcher.py:215-283) and
`search_by_projection_f_f` (ORBMatcher.py:291-393) -- with the Hamming distances taken from one GPU all-pairs matrix
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
    D = hamming_matrix(kf.mDescriptors, frame.mDescriptors) if frame.N and len(mps_kf) else None
    n_matches = 0
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    fv_kf, fv_f = kf.mFeatVec, frame.mFeatVec
    it_kf, it_f = iter(fv_kf), iter(fv_f)
    try:
        node_kf, node_f = next(it_kf), next(it_f)
        while True:
            if node_kf == node_f:
                idx_f_list = fv_f[node_f]
                idx_f_arr = np.asarray(idx_f_list, np.intp)
                for i_kf in fv_kf[node_kf]:
                    mp = mps_kf[i_kf]
                    if not mp or mp.is_bad():
                        continue
                    dists = D[i_kf, idx_f_arr].tolist()          # distances of this keyframe feature to the node's frame features
                    best1, best_f, best2 = 256, -1, 256
                    for i_f, d in zip(idx_f_list, dists):
                        if matches[i_f]:
                            continue
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
    D = hamming_matrix(kf1.mDescriptors, kf2.mDescriptors) if len(mps1) and len(mps2) else None
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    n_matches = 0
    it1, it2 = iter(kf1.mFeatVec.items()), iter(kf2.mFeatVec.items())
    try:
        f1, f2 = next(it1), next(it2)
        while True:
            if f1[0] == f2[0]:
                idx2_arr = np.asarray(f2[1], np.intp)
                for i1 in f1[1]:
                    mp1 = mps1[i1]
                    if not mp1 or mp1.is_bad():
                        continue
                    dists = D[i1, idx2_arr].tolist()
                    best1, best_i2, best2 = 256, -1, 256
                    for i2, d in zip(f2[1], dists):
                        mp2 = mps2[i2]
                        if matched2[i2] or not mp2 or mp2.is_bad():
                            continue
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


def _descriptor_rows(map_points, wanted):
    """uint8[n, 32]: row i = map_points[i].get_descriptor() where wanted[i], zeros elsewhere."""
    out = np.zeros((len(map_points), 32), np.uint8)
    for i, (mp, w) in enumerate(zip(map_points, wanted)):
        if w:
            out[i] = mp.get_descriptor()
    return out


def search_by_projection_f_f(self, current_frame, last_frame, th):
    """ORBMatcher.search_by_projection_f_f, ORBMatcher.py:291-393 (TrackWithMotionModel)."""
    n_matches = 0
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    Rcw = current_frame.mTcw[:3, :3]
    tcw = current_frame.mTcw[:3, 3:4]
    twc = -Rcw.T @ tcw
    Rlw = last_frame.mTcw[:3, :3]
    tlw = last_frame.mTcw[:3, 3:4]
    tlc = Rlw @ twc + tlw
    b_forward = tlc[2] > current_frame.mb
    b_backward = -tlc[2] > current_frame.mb
    usable = [bool(last_frame.mvpMapPoints[i]) and not last_frame.mvbOutlier[i] for i in range(last_frame.N)]
    D = hamming_matrix(_descriptor_rows(last_frame.mvpMapPoints, usable), current_frame.mDescriptors) if any(usable) and current_frame.N else None
    cur_mps, cur_ur = current_frame.mvpMapPoints, current_frame.mvuRight
    for i in range(last_frame.N):
        if not usable[i]:
            continue
        pMP = last_frame.mvpMapPoints[i]
        x3Dc = Rcw @ pMP.get_world_pos() + tcw
        xc, yc, zc = x3Dc[0][0], x3Dc[1][0], x3Dc[2][0]
        invzc = 1.0 / zc
        if invzc < 0:
            continue
        u = current_frame.fx * xc * invzc + current_frame.cx
        v = current_frame.fy * yc * invzc + current_frame.cy
        if u < current_frame.mnMinX or u > current_frame.mnMaxX:
            continue
        if v < current_frame.mnMinY or v > current_frame.mnMaxY:
            continue
        octave = last_frame.mvKeys[i].octave
        radius = th * current_frame.mvScaleFactors[octave]
        if b_forward:
            cand = current_frame.get_features_in_area(u, v, radius, octave, -1)
        elif b_backward:
            cand = current_frame.get_features_in_area(u, v, radius, 0, octave)
        else:
            cand = current_frame.get_features_in_area(u, v, radius, octave - 1, octave + 1)
        if not cand:
            continue
        dists = D[i, np.asarray(cand, np.intp)].tolist()
        best_dist, best_idx2 = 256, -1
        for i2, dist in zip(cand, dists):
            if cur_mps[i2]:
                if cur_mps[i2].observations() > 0:
                    continue
            if cur_ur[i2] > 0:
                ur = u - current_frame.mbf * invzc
                if abs(ur - cur_ur[i2]) > radius:
                    continue
            if dist < best_dist:
                best_dist, best_idx2 = dist, i2
        if best_dist <= TH_HIGH:
            cur_mps[best_idx2] = pMP
            n_matches += 1
            if self.mbCheckOrientation:
                b = _rot_bin(last_frame.mvKeysUn[i].angle - current_frame.mvKeysUn[best_idx2].angle)
                assert 0 <= b < HISTO_LENGTH
                rot_hist[b].append(best_idx2)
    if self.mbCheckOrientation:
        keep = _three_maxima(rot_hist)
        for i in range(HISTO_LENGTH):
            if i not in keep:
                for idx in rot_hist[i]:
                    cur_mps[idx] = None
                    n_matches -= 1
    return n_matches


def search_by_projection_f_p(self, frame, vp_map_points, th):
    """ORBMatcher.search_by_projection_f_p, ORBMatcher.py:215-283 (SearchLocalPoints)."""
    n_matches = 0
    b_factor = th != 1.0
    usable = [bool(mp.mbTrackInView) and not mp.is_bad() for mp in vp_map_points]
    D = hamming_matrix(_descriptor_rows(vp_map_points, usable), frame.mDescriptors) if any(usable) and frame.N else None
    for k, pMP in enumerate(vp_map_points):
        if not usable[k]:
            continue
        level = pMP.mnTrackScaleLevel
        r = 2.5 if pMP.mTrackViewCos > 0.998 else 4.0          # radius_by_viewing_cos, ORBMatcher.py:285-289
        if b_factor:
            r *= th
        cand = frame.get_features_in_area(pMP.mTrackProjX, pMP.mTrackProjY, r * frame.mvScaleFactors[level], level - 1, level)
        if not cand:
            continue
        dists = D[k, np.asarray(cand, np.intp)].tolist()
        best_dist, best_level, best_dist2, best_level2, best_idx = 256, -1, 256, -1, -1
        for idx, dist in zip(cand, dists):
            if frame.mvpMapPoints[idx]:
                if frame.mvpMapPoints[idx].observations() > 0:
                    continue
            if frame.mvuRight[idx] > 0:
                if abs(pMP.mTrackProjXR - frame.mvuRight[idx]) > r * frame.mvScaleFactors[level]:
                    continue
            if dist < best_dist:
                best_dist2, best_dist = best_dist, dist
                best_level2, best_level = best_level, frame.mvKeysUn[idx].octave
                best_idx = idx
            elif dist < best_dist2:
                best_level2 = frame.mvKeysUn[idx].octave
                best_dist2 = dist
        if best_dist <= TH_HIGH:
            if best_level == best_level2 and best_dist > self.mfNNratio * best_dist2:
                continue
            frame.mvpMapPoints[best_idx] = pMP
            n_matches += 1
    return n_matches


def install_matcher(matcher_cls):
    """Patch the reference's ORBMatcher class (BoW searches + the two frame projection searches); returns the originals."""
    names = ("search_by_BoW_kf_f", "search_by_BoW_kf_kf", "search_by_projection_f_f", "search_by_projection_f_p")
    orig = {n: getattr(matcher_cls, n) for n in names if hasattr(matcher_cls, n)}
    matcher_cls.search_by_BoW_kf_f = search_by_BoW_kf_f
    matcher_cls.search_by_BoW_kf_kf = search_by_BoW_kf_kf
    matcher_cls.search_by_projection_f_f = search_by_projection_f_f
    matcher_cls.search_by_projection_f_p = search_by_projection_f_p
    return orig
