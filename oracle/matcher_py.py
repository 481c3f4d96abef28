"""TEST INFRASTRUCTURE -- Python restatement of ORBMatcher.search_by_BoW_kf_f / search_by_BoW_kf_kf
(reference ORBMatcher.py:12-213) with the reference's own per-pair Python Hamming distance.  Pinned against vectors produced
by the reference's ORBMatcher class itself (tests/golden/make_golden.py::matcher_case -> matcher_small.npz)."""
import types

import numpy as np

TH_LOW, HISTO_LENGTH = 50, 30


def distance(a, b):                      # ORBMatcher.py:12-14
    return sum(bin(int(v)).count("1") for v in np.bitwise_xor(a, b))


def _three(rot_hist):
    return np.argsort([len(h) for h in rot_hist])[::-1][:3]


def _bin(rot):
    if rot < 0.0:
        rot += 360.0
    b = round(rot * (1.0 / HISTO_LENGTH))
    return 0 if b == HISTO_LENGTH else b


def _merge(fa, fb):
    """Nodes present in both ordered feature vectors, in order (the reference's two-iterator merge)."""
    ia, ib = iter(fa), iter(fb)
    try:
        a, b = next(ia), next(ib)
        while True:
            if a == b:
                yield a
                a, b = next(ia), next(ib)
            elif a < b:
                a = next(ia)
            else:
                b = next(ib)
    except StopIteration:
        return


def bow_kf_f(kf, frame, nnratio=1, check_ori=True):
    mps = kf.get_map_point_matches()
    out = [None] * frame.N
    n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    for node in _merge(kf.mFeatVec, frame.mFeatVec):
        for ik in kf.mFeatVec[node]:
            mp = mps[ik]
            if not mp or mp.is_bad():
                continue
            b1, bf, b2 = 256, -1, 256
            for jf in frame.mFeatVec[node]:
                if out[jf]:
                    continue
                d = distance(kf.mDescriptors[ik], frame.mDescriptors[jf])
                if d < b1:
                    b2, b1, bf = b1, d, jf
                elif d < b2:
                    b2 = d
            if b1 <= TH_LOW and float(b1) < nnratio * float(b2):
                out[bf] = mp
                if check_ori:
                    hist[_bin(kf.mvKeysUn[ik].angle - frame.mvKeys[bf].angle)].append(bf)
                n += 1
    if check_ori:
        keep = _three(hist)
        for i in range(HISTO_LENGTH):
            if i not in keep:
                for idx in hist[i]:
                    out[idx] = None
                    n -= 1
    return n, out


def bow_kf_kf(k1, k2, nnratio=1, check_ori=True):
    m1, m2 = k1.get_map_point_matches(), k2.get_map_point_matches()
    out = [None] * len(m1)
    used = [False] * len(m2)
    n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    for node in _merge(k1.mFeatVec, k2.mFeatVec):
        for i1 in k1.mFeatVec[node]:
            if not m1[i1] or m1[i1].is_bad():
                continue
            b1, bi, b2 = 256, -1, 256
            for i2 in k2.mFeatVec[node]:
                if used[i2] or not m2[i2] or m2[i2].is_bad():
                    continue
                d = distance(k1.mDescriptors[i1], k2.mDescriptors[i2])
                if d < b1:
                    b2, b1, bi = b1, d, i2
                elif d < b2:
                    b2 = d
            if b1 < TH_LOW and b1 < nnratio * b2:
                out[i1] = m2[bi]
                used[bi] = True
                if check_ori:
                    hist[_bin(k1.mvKeysUn[i1].angle - k2.mvKeysUn[bi].angle)].append(i1)
                n += 1
    if check_ori:
        keep = _three(hist)
        for i in range(HISTO_LENGTH):
            if i not in keep:
                for idx in hist[i]:
                    out[idx] = None
                    n -= 1
    return n, out


# ---------------------------------------------------------------- synthetic scene for the tests / golden vectors
class FakeMapPoint:
    def __init__(self, uid, bad):
        self.uid, self._bad = uid, bad

    def is_bad(self):
        return self._bad


def make_case(seed=3, n=320, n_nodes=6):
    """Two 'keyframes' observing the same scene: descriptors of B are noisy copies of A's (random permutation), both carry a
    FeatureVector over `n_nodes` vocabulary nodes, some features have no map point, some map points are bad."""
    from collections import OrderedDict
    rng = np.random.default_rng(seed)
    dA = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    twins = rng.choice(n, n // 3, replace=False)        # near-duplicate descriptors: the best / second-best ratio test bites
    dA[twins] = dA[(twins + 1) % n] ^ np.packbits(rng.random((len(twins), 256)) < 0.04, axis=1, bitorder="little")
    perm = rng.permutation(n)
    flips = np.packbits(rng.random((n, 256)) < rng.choice([0.03, 0.12, 0.3], size=(n, 1)), axis=1, bitorder="little")
    dB = (dA ^ flips)[perm]
    node_of_A = rng.integers(0, n_nodes, n)
    node_of_B = np.where(rng.random(n) < 0.9, node_of_A[perm], rng.integers(0, n_nodes, n))

    def featvec(nodes):
        fv = {}
        for i, nd in enumerate(nodes.tolist()):
            fv.setdefault(nd * 7 + 3, []).append(i)
        return OrderedDict(sorted(fv.items()))
    angA = rng.uniform(0, 360, n).astype(np.float32)
    angB = ((angA[perm] + rng.choice([0.0, 40.0, 80.0, 120.0, 200.0], size=n, p=[0.5, 0.2, 0.15, 0.1, 0.05]) + rng.normal(0, 2, n)) % 360).astype(np.float32)

    def side(desc, nodes, ang, off):
        mps = [None if rng.random() < 0.15 else FakeMapPoint(off + i, bool(rng.random() < 0.05)) for i in range(n)]
        keys = [types.SimpleNamespace(angle=float(a)) for a in ang]
        return types.SimpleNamespace(mDescriptors=desc, mFeatVec=featvec(nodes), mvKeysUn=keys, mvKeys=keys, N=n,
                                     get_map_point_matches=lambda m=mps: m, mps=mps)
    return side(dA, node_of_A, angA, 0), side(dB, node_of_B, angB, 10000)


# ================================================================ projection searches (ORBMatcher.py:215-393)
def assign_grid(frame):
    """Frame.assign_features_to_grid + pos_in_grid, Frame.py:143-159."""
    grid = [[[] for _ in range(frame.FRAME_GRID_ROWS)] for _ in range(frame.FRAME_GRID_COLS)]
    pts = np.array([[k.pt[0], k.pt[1]] for k in frame.mvKeys])
    px = np.round((pts[:, 0] - frame.mnMinX) * frame.mfGridElementWidthInv).astype(int)
    py = np.round((pts[:, 1] - frame.mnMinY) * frame.mfGridElementHeightInv).astype(int)
    for i in range(frame.N):
        if 0 <= px[i] < frame.FRAME_GRID_COLS and 0 <= py[i] < frame.FRAME_GRID_ROWS:
            grid[px[i]][py[i]].append(i)
    return grid


def features_in_area(frame, x, y, r, min_level, max_level):
    """Frame.get_features_in_area, Frame.py:373-416."""
    out = []
    c0 = max(0, int((x - frame.mnMinX - r) * frame.mfGridElementWidthInv))
    if c0 >= frame.FRAME_GRID_COLS:
        return out
    c1 = min(frame.FRAME_GRID_COLS - 1, int((x - frame.mnMinX + r) * frame.mfGridElementWidthInv))
    if c1 < 0:
        return out
    r0 = max(0, int((y - frame.mnMinY - r) * frame.mfGridElementHeightInv))
    if r0 >= frame.FRAME_GRID_ROWS:
        return out
    r1 = min(frame.FRAME_GRID_ROWS - 1, int((y - frame.mnMinY + r) * frame.mfGridElementHeightInv))
    if r1 < 0:
        return out
    check = (min_level > 0) or (max_level >= 0)
    for ix in range(c0, c1 + 1):
        for iy in range(r0, r1 + 1):
            for g in frame.mGrid[ix][iy]:
                kp = frame.mvKeysUn[g]
                if check:
                    if kp.octave < min_level:
                        continue
                    if max_level >= 0 and kp.octave > max_level:
                        continue
                if abs(kp.pt[0] - x) < r and abs(kp.pt[1] - y) < r:
                    out.append(g)
    return out


def projection_f_f(cur, last, th, check_ori=True):
    n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    Rcw, tcw = cur.mTcw[:3, :3], cur.mTcw[:3, 3:4]
    tlc = last.mTcw[:3, :3] @ (-Rcw.T @ tcw) + last.mTcw[:3, 3:4]
    fwd, bwd = tlc[2] > cur.mb, -tlc[2] > cur.mb
    for i in range(last.N):
        mp = last.mvpMapPoints[i]
        if not mp or last.mvbOutlier[i]:
            continue
        pc = Rcw @ mp.get_world_pos() + tcw
        xc, yc, zc = pc[0][0], pc[1][0], pc[2][0]
        iz = 1.0 / zc
        if iz < 0:
            continue
        u, v = cur.fx * xc * iz + cur.cx, cur.fy * yc * iz + cur.cy
        if u < cur.mnMinX or u > cur.mnMaxX or v < cur.mnMinY or v > cur.mnMaxY:
            continue
        o = last.mvKeys[i].octave
        rad = th * cur.mvScaleFactors[o]
        lo, hi = (o, -1) if fwd else ((0, o) if bwd else (o - 1, o + 1))
        cand = cur.get_features_in_area(u, v, rad, lo, hi)
        if not cand:
            continue
        d_mp = mp.get_descriptor()
        best, bi = 256, -1
        for j in cand:
            if cur.mvpMapPoints[j] and cur.mvpMapPoints[j].observations() > 0:
                continue
            if cur.mvuRight[j] > 0 and abs(u - cur.mbf * iz - cur.mvuRight[j]) > rad:
                continue
            d = distance(d_mp, cur.mDescriptors[j])
            if d < best:
                best, bi = d, j
        if best <= 100:
            cur.mvpMapPoints[bi] = mp
            n += 1
            if check_ori:
                hist[_bin(last.mvKeysUn[i].angle - cur.mvKeysUn[bi].angle)].append(bi)
    if check_ori:
        keep = _three(hist)
        for i in range(HISTO_LENGTH):
            if i not in keep:
                for j in hist[i]:
                    cur.mvpMapPoints[j] = None
                    n -= 1
    return n


def projection_f_p(frame, points, th, nnratio=1):
    n = 0
    for mp in points:
        if not mp.mbTrackInView or mp.is_bad():
            continue
        lvl = mp.mnTrackScaleLevel
        r = (2.5 if mp.mTrackViewCos > 0.998 else 4.0)
        if th != 1.0:
            r *= th
        cand = frame.get_features_in_area(mp.mTrackProjX, mp.mTrackProjY, r * frame.mvScaleFactors[lvl], lvl - 1, lvl)
        if not cand:
            continue
        d_mp = mp.get_descriptor()
        b1, l1, b2, l2, bi = 256, -1, 256, -1, -1
        for j in cand:
            if frame.mvpMapPoints[j] and frame.mvpMapPoints[j].observations() > 0:
                continue
            if frame.mvuRight[j] > 0 and abs(mp.mTrackProjXR - frame.mvuRight[j]) > r * frame.mvScaleFactors[lvl]:
                continue
            d = distance(d_mp, frame.mDescriptors[j])
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, frame.mvKeysUn[j].octave, j
            elif d < b2:
                l2, b2 = frame.mvKeysUn[j].octave, d
        if b1 <= 100:
            if l1 == l2 and b1 > nnratio * b2:
                continue
            frame.mvpMapPoints[bi] = mp
            n += 1
    return n


class FakeTrackedPoint(FakeMapPoint):
    def __init__(self, uid, bad, pos, desc, nobs):
        super().__init__(uid, bad)
        self._pos, self._desc, self._nobs = pos, desc, nobs

    def get_world_pos(self):
        return self._pos.copy()

    def get_descriptor(self):
        return self._desc.copy()

    def observations(self):
        return self._nobs


def make_projection_case(seed=4, n=600, W=1241, H=376, get_area=None, assign=None, motion=(0.05, 0.0, 0.3)):
    """A 'last' and a 'current' frame seeing the same 3-D points after a small camera translation `motion`.
    get_area / assign: the functions to bind as Frame.get_features_in_area / Frame.assign_features_to_grid
    (the reference's own when generating golden vectors, the restatements above in the tests)."""
    rng = np.random.default_rng(seed)
    fx = fy = 718.856
    cx, cy, mbf = 607.1928, 185.2157, 386.1448
    sf = [float(np.float32(1.2) ** l) for l in range(8)]

    def frame(Tcw, pts_w, desc, octv, ang):
        pc = (Tcw[:3, :3] @ pts_w.T + Tcw[:3, 3:4]).T
        u = fx * pc[:, 0] / pc[:, 2] + cx + rng.normal(0, 0.7, len(pc))
        v = fy * pc[:, 1] / pc[:, 2] + cy + rng.normal(0, 0.7, len(pc))
        keys = [types.SimpleNamespace(pt=(float(np.float32(a)), float(np.float32(b))), octave=int(o), angle=float(g))
                for a, b, o, g in zip(u, v, octv, ang)]
        f = types.SimpleNamespace(N=len(keys), mvKeys=keys, mvKeysUn=keys, mDescriptors=desc, mTcw=Tcw.astype(np.float32), fx=fx, fy=fy,
                                  cx=cx, cy=cy, mb=np.float32(mbf) / np.float32(fx), mbf=mbf, mnMinX=0.0, mnMaxX=float(W), mnMinY=0.0,
                                  mnMaxY=float(H), FRAME_GRID_ROWS=48, FRAME_GRID_COLS=64, mfGridElementWidthInv=64.0 / W,
                                  mfGridElementHeightInv=48.0 / H, mvScaleFactors=sf, mvpMapPoints=[None] * len(keys),
                                  mvbOutlier=[False] * len(keys))
        disp = mbf / pc[:, 2]
        f.mvuRight = [float(np.float32(a - d)) if rng.random() < 0.6 else -1 for a, d in zip(u, disp)]
        f.pos_in_grid = types.MethodType(lambda self, kps: None, f)
        f.mGrid = (assign or assign_grid)(f)
        f.get_features_in_area = types.MethodType(get_area or features_in_area, f)
        return f
    pts = np.stack([rng.uniform(-12, 12, n), rng.uniform(-3, 3, n), rng.uniform(6, 40, n)], 1)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    octv = rng.integers(0, 8, n)
    ang = rng.uniform(0, 360, n)
    T_last = np.eye(4)
    T_cur = np.eye(4)
    T_cur[:3, 3] = motion
    last = frame(T_last, pts, desc, octv, ang)
    noisy = desc ^ np.packbits(rng.random((n, 256)) < rng.choice([0.02, 0.1, 0.35], size=(n, 1)), axis=1, bitorder="little")
    perm = rng.permutation(n)
    cur = frame(T_cur, pts[perm], noisy[perm], np.clip(octv[perm] + rng.integers(-1, 2, n), 0, 7),
                (ang[perm] + rng.choice([0.0, 30.0, 170.0], size=n, p=[0.75, 0.15, 0.1]) + rng.normal(0, 2, n)) % 360)
    for i in range(n):      # map points of the last frame; a few features without a point, a few outliers
        if rng.random() < 0.85:
            last.mvpMapPoints[i] = FakeTrackedPoint(i, False, pts[i].reshape(3, 1).astype(np.float32), desc[i], int(rng.integers(1, 6)))
        last.mvbOutlier[i] = bool(rng.random() < 0.05)
    for j in range(0, n, 9):  # the current frame already holds some points (with / without observations)
        cur.mvpMapPoints[j] = FakeTrackedPoint(50000 + j, False, np.zeros((3, 1), np.float32), cur.mDescriptors[j], int(rng.integers(0, 2)))
    # local map points for search_by_projection_f_p: projected into `cur` with predicted levels
    local = []
    pcur = (T_cur[:3, :3] @ pts.T + T_cur[:3, 3:4]).T
    for i in range(n):
        mp = FakeTrackedPoint(90000 + i, bool(rng.random() < 0.04), pts[i].reshape(3, 1).astype(np.float32), desc[i], int(rng.integers(1, 5)))
        mp.mbTrackInView = bool(rng.random() < 0.9)
        mp.mnTrackScaleLevel = int(octv[i])
        mp.mTrackViewCos = float(rng.choice([0.9995, 0.98]))
        mp.mTrackProjX = float(fx * pcur[i, 0] / pcur[i, 2] + cx)
        mp.mTrackProjY = float(fy * pcur[i, 1] / pcur[i, 2] + cy)
        mp.mTrackProjXR = float(mp.mTrackProjX - mbf / pcur[i, 2])
        local.append(mp)
    return cur, last, local
