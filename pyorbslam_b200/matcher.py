"""the tracking-side searches of the reference's ORBMatcher -- `search_by_BoW_kf_f`,
`search_by_BoW_kf_kf` (ORBMatcher.py:21-213), `search_by_projection_f_p` (ORBMatcher.py:215-283) and
`search_by_projection_f_f` (ORBMatcher.py:291-393) -- with the Hamming distances taken from the GPU instead of one Python
`bin().count` call per candidate pair (ORBMatcher.py:12-14): an all-pairs matrix (`b200orb_hamming_matrix`) for the BoW searches,
and for the projection searches one call that also answers every `Frame.get_features_in_area` query (`b200orb_area_hamming`)
followed by the order-dependent selection behind the same C ABI (`b200orb_greedy_project_ff / _fp`).

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


# ---------------------------------------------------------------- projection searches
def _elem_float_dtype(values):
    """dtype of the NumPy floating scalars in `values` (None when there are only Python numbers, which NumPy >= 2 treats as weak)."""
    dt = None
    for v in values:
        if isinstance(v, np.floating):
            dt = v.dtype if dt is None else np.result_type(dt, v.dtype)
    return dt


def _frame_arrays(frame):
    """Array views of a frame's immutable-after-construction data (undistorted keypoints, grid, mvuRight), built once per frame and
    kept on the frame object: TrackWithMotionModel and SearchLocalPoints query the same current frame several times.  The key is the
    identity of the three list objects (Frame.__init__ assigns them once, Frame.py:57-70)."""
    key = (id(frame.mvKeysUn), id(frame.mGrid), id(frame.mvuRight), frame.N)
    c = getattr(frame, "_b200orb_arrays", None)
    if c is None or c["key"] != key:
        cs, ci = _grid_csr(frame)
        ur_list = frame.mvuRight
        c = {"key": key,
             "kxy": np.array([k.pt for k in frame.mvKeysUn], np.float32).reshape(-1, 2),
             "koct": np.array([k.octave for k in frame.mvKeysUn], np.int32),
             "cs": cs, "ci": ci,
             "ur_el": _elem_float_dtype(ur_list),
             "ur": np.array([float(v) for v in ur_list], np.float64)}
        try:
            frame._b200orb_arrays = c
        except AttributeError:      # frames with __slots__: no cache
            pass
    return c


def _grid_csr(frame):
    """frame.mGrid (Frame.assign_features_to_grid, Frame.py:153-159) as CSR over cell ix * rows + iy."""
    cols, rows = frame.FRAME_GRID_COLS, frame.FRAME_GRID_ROWS
    start = np.zeros(cols * rows + 1, np.int32)
    flat = []
    k = 0
    for ix in range(cols):
        col = frame.mGrid[ix]
        for iy in range(rows):
            cell = col[iy]
            if cell:
                flat.extend(cell)
            k += 1
            start[k] = len(flat)
    return start, np.asarray(flat, np.int32).reshape(-1)


def _cell_ranges(frame, x, y, r):
    """n_min_cell_x .. n_max_cell_y of Frame.get_features_in_area (Frame.py:376-390) for arrays of queries, in the arithmetic the
    reference's scalar expressions run in: x / y keep their dtype, the frame's Python floats and the radius are cast to it."""
    dt = x.dtype
    r = r.astype(dt)
    minx, miny = dt.type(frame.mnMinX), dt.type(frame.mnMinY)
    iw, ih = dt.type(frame.mfGridElementWidthInv), dt.type(frame.mfGridElementHeightInv)
    cols, rows = frame.FRAME_GRID_COLS, frame.FRAME_GRID_ROWS
    with np.errstate(invalid="ignore", over="ignore"):
        c0 = np.maximum(0, np.trunc((x - minx - r) * iw)).astype(np.int64)
        c1 = np.minimum(cols - 1, np.trunc((x - minx + r) * iw)).astype(np.int64)
        r0 = np.maximum(0, np.trunc((y - miny - r) * ih)).astype(np.int64)
        r1 = np.minimum(rows - 1, np.trunc((y - miny + r) * ih)).astype(np.int64)
    empty = (c0 >= cols) | (c1 < 0) | (r0 >= rows) | (r1 < 0)          # the four early returns
    out = np.stack([c0, c1, r0, r1], 1).astype(np.int32)
    out[empty] = (1, 0, 0, 0)
    return out


def _area_hamming(frame, qx, qy, qr, qlvl, qdesc, device=0):
    """Candidate lists of Frame.get_features_in_area for every query + Hamming distance to each candidate (one GPU call)."""
    M = len(qx)
    start = np.zeros(M + 1, np.int32)
    if M == 0 or frame.N == 0:
        return start, np.zeros(0, np.int32), np.zeros(0, np.int32)
    fa = _frame_arrays(frame)
    kxy, koct, cs, ci = fa["kxy"], fa["koct"], fa["cs"], fa["ci"]
    kdesc = np.ascontiguousarray(frame.mDescriptors, np.uint8).reshape(-1, 32)
    f32 = qx.dtype == np.float32
    qcell = _cell_ranges(frame, qx, qy, qr)
    qxyr = np.ascontiguousarray(np.stack([qx.astype(np.float64), qy.astype(np.float64),
                                          qr.astype(qx.dtype).astype(np.float64)], 1))
    qlvl = np.ascontiguousarray(qlvl, np.int32)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    l = _lib.lib()
    l.b200orb_area_hamming.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 2 + \
                                      [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
    cap = max(64 * M, 1024)
    total = C.c_int32(0)
    while True:
        idx = np.empty(cap, np.int32)
        dist = np.empty(cap, np.int32)
        rc = l.b200orb_area_hamming(int(device), int(f32), int(frame.N), kxy.ctypes.data, koct.ctypes.data, kdesc.ctypes.data,
                                    int(frame.FRAME_GRID_COLS), int(frame.FRAME_GRID_ROWS), cs.ctypes.data, ci.ctypes.data, M,
                                    qxyr.ctypes.data, qlvl.ctypes.data, qcell.ctypes.data, qdesc.ctypes.data, start.ctypes.data,
                                    idx.ctypes.data, dist.ctypes.data, cap, C.byref(total))
        if rc != 0 and total.value > cap:
            cap = total.value
            continue
        _lib.check(rc)
        return start, idx[:total.value], dist[:total.value]


def _right_ok(frame, start, idx, q_xr, q_rad):
    """Per candidate: False when the right-image check rejects it (ORBMatcher.py:246-249 / 345-349), evaluated in the dtype the
    reference's scalar expression `abs(xr - mvuRight[i]) > radius` runs in."""
    n = len(idx)
    if n == 0:
        return np.ones(0, np.uint8)
    fa = _frame_arrays(frame)
    el, ur = fa["ur_el"], fa["ur"]
    dt = q_xr.dtype if el is None else np.result_type(q_xr.dtype, el)
    has_r = ur[idx] > 0
    per_q = np.repeat(np.arange(len(start) - 1), np.diff(start))
    er = np.abs(q_xr.astype(dt)[per_q] - ur.astype(dt)[idx])
    bad = has_r & (er > q_rad.astype(dt)[per_q])
    return (~bad).astype(np.uint8)


def _occupied(frame):
    return np.array([1 if (mp and mp.observations() > 0) else 0 for mp in frame.mvpMapPoints], np.uint8)


def _homogeneous(vals):
    kinds = {type(v) for v in vals}
    return len(kinds) <= 1 or all(not isinstance(v, np.floating) for v in vals)


def search_by_projection_f_f(self, current_frame, last_frame, th):
    """ORBMatcher.search_by_projection_f_f, ORBMatcher.py:291-393 (TrackWithMotionModel).
    The projections of all map points are one stacked 3x3 @ 3x1 matmul plus elementwise expressions of the same dtype as the
    reference's scalar ones (bit-identical, self-checked per call); all window queries and descriptor distances are one GPU call
    (b200orb_area_hamming), the order-dependent selection one host call (b200orb_greedy_project_ff)."""
    Rcw = current_frame.mTcw[:3, :3]
    tcw = current_frame.mTcw[:3, 3:4]
    twc = -Rcw.T @ tcw
    Rlw = last_frame.mTcw[:3, :3]
    tlw = last_frame.mTcw[:3, 3:4]
    tlc = Rlw @ twc + tlw
    b_forward = tlc[2] > current_frame.mb
    b_backward = -tlc[2] > current_frame.mb
    fx, fy, cx, cy, mbf = current_frame.fx, current_frame.fy, current_frame.cx, current_frame.cy, current_frame.mbf
    x0, x1, y0, y1 = current_frame.mnMinX, current_frame.mnMaxX, current_frame.mnMinY, current_frame.mnMaxY
    sf = current_frame.mvScaleFactors
    cand = [i for i in range(last_frame.N) if last_frame.mvpMapPoints[i] and not last_frame.mvbOutlier[i]]      # ORBMatcher.py:299-303
    if not cand or current_frame.N == 0:
        return 0
    mps = [last_frame.mvpMapPoints[i] for i in cand]
    Pw = np.stack([np.asarray(mp.get_world_pos()) for mp in mps])             # [n, 3, 1]
    # x3Dc = Rcw @ pos + tcw for all points as ONE stacked matmul: NumPy runs the same 3x3 @ 3x1 routine per slice, so the results
    # are the per-point ones bit for bit (an einsum or a hand-written sum is NOT: different association / contraction).  Checked on
    # the first points of every call; if this NumPy build ever disagrees, the per-point expression of the reference is used.
    X = np.matmul(Rcw[None], Pw) + tcw[None]
    nchk = min(len(cand), 4)
    if not all(np.array_equal(X[j], Rcw @ Pw[j] + tcw) for j in range(nchk)):
        X = np.stack([Rcw @ p + tcw for p in Pw])
    xc, yc, zc = X[:, 0, 0], X[:, 1, 0], X[:, 2, 0]
    with np.errstate(divide="ignore", invalid="ignore"):
        invzc = 1.0 / zc                                                       # ORBMatcher.py:313
        u = fx * xc * invzc + cx
        v = fy * yc * invzc + cy
        keep = ~(invzc < 0) & ~((u < x0) | (u > x1)) & ~((v < y0) | (v > y1))  # the three `continue`s, ORBMatcher.py:314-323
    sel = np.nonzero(keep)[0]
    M = len(sel)
    if M == 0:
        return 0
    qi = [cand[j] for j in sel.tolist()]
    qmp = [mps[j] for j in sel.tolist()]
    octs = [last_frame.mvKeys[i].octave for i in qi]
    qu, qv = np.ascontiguousarray(u[sel]), np.ascontiguousarray(v[sel])
    qxr = qu - mbf * invzc[sel]                                                # ORBMatcher.py:347 (same dtype as the scalar expression)
    qrad = np.array([th * sf[o] for o in octs], np.float64)
    qlvl = [(o, -1) if b_forward else ((0, o) if b_backward else (o - 1, o + 1)) for o in octs]
    if qu.dtype != qv.dtype or qu.dtype not in (np.float32, np.float64):
        raise TypeError("projected coordinates must be float32 or float64")
    qdesc = np.stack([np.asarray(mp.get_descriptor(), np.uint8).reshape(32) for mp in qmp])
    start, idx, dist = _area_hamming(current_frame, qu, qv, qrad, np.array(qlvl, np.int32), qdesc)
    ok = _right_ok(current_frame, start, idx, qxr, qrad)
    occ = _occupied(current_frame)
    marks = np.array([1 if mp.observations() > 0 else 0 for mp in qmp], np.uint8)
    best = np.empty(M, np.int32)
    l = _lib.lib()
    l.b200orb_greedy_project_ff.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    _lib.check(l.b200orb_greedy_project_ff(M, start.ctypes.data, idx.ctypes.data, dist.ctypes.data, ok.ctypes.data, int(current_frame.N),
                                           occ.ctypes.data, marks.ctypes.data, TH_HIGH, best.ctypes.data))
    n_matches = 0
    rot_hist = [[] for _ in range(HISTO_LENGTH)]
    cur_mps = current_frame.mvpMapPoints
    for q in np.nonzero(best >= 0)[0].tolist():
        b2 = int(best[q])
        cur_mps[b2] = qmp[q]
        n_matches += 1
        if self.mbCheckOrientation:
            b = _rot_bin(last_frame.mvKeysUn[qi[q]].angle - current_frame.mvKeysUn[b2].angle)
            assert 0 <= b < HISTO_LENGTH
            rot_hist[b].append(b2)
    if self.mbCheckOrientation:
        keep = _three_maxima(rot_hist)
        for i in range(HISTO_LENGTH):
            if i not in keep:
                for k in rot_hist[i]:
                    cur_mps[k] = None
                    n_matches -= 1
    return n_matches


def search_by_projection_f_p(self, frame, vp_map_points, th):
    """ORBMatcher.search_by_projection_f_p, ORBMatcher.py:215-283 (SearchLocalPoints); same split as search_by_projection_f_f."""
    b_factor = th != 1.0
    qmp, qx, qy, qxr, qrad, qlvl = [], [], [], [], [], []
    sf = frame.mvScaleFactors
    for pMP in vp_map_points:
        if not pMP.mbTrackInView:
            continue
        if pMP.is_bad():
            continue
        level = pMP.mnTrackScaleLevel
        r = 2.5 if pMP.mTrackViewCos > 0.998 else 4.0          # radius_by_viewing_cos, ORBMatcher.py:285-289
        if b_factor:
            r *= th
        qmp.append(pMP); qx.append(pMP.mTrackProjX); qy.append(pMP.mTrackProjY); qxr.append(pMP.mTrackProjXR)
        qrad.append(r * sf[level]); qlvl.append((level - 1, level))
    M = len(qmp)
    if M == 0 or frame.N == 0:
        return 0
    if not (_homogeneous(qx) and _homogeneous(qy) and _homogeneous(qxr)):
        raise TypeError("projected coordinates of mixed scalar types")
    qx, qy, qxr = np.array(qx), np.array(qy), np.array(qxr)
    if qx.dtype != qy.dtype or qx.dtype not in (np.float32, np.float64):
        raise TypeError("projected coordinates must be float32 or float64 scalars")
    qrad = np.array(qrad, np.float64)
    qdesc = np.stack([np.asarray(mp.get_descriptor(), np.uint8).reshape(32) for mp in qmp])
    start, idx, dist = _area_hamming(frame, qx, qy, qrad, np.array(qlvl, np.int32), qdesc)
    ok = _right_ok(frame, start, idx, qxr, qrad)
    occ = _occupied(frame)
    marks = np.array([1 if mp.observations() > 0 else 0 for mp in qmp], np.uint8)
    koct = _frame_arrays(frame)["koct"]
    best = np.empty(M, np.int32)
    l = _lib.lib()
    l.b200orb_greedy_project_fp.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                                                            C.c_void_p]
    _lib.check(l.b200orb_greedy_project_fp(M, start.ctypes.data, idx.ctypes.data, dist.ctypes.data, ok.ctypes.data, int(frame.N),
                                           occ.ctypes.data, marks.ctypes.data, koct.ctypes.data, TH_HIGH, float(self.mfNNratio),
                                           best.ctypes.data))
    n_matches = 0
    for q in np.nonzero(best >= 0)[0].tolist():
        frame.mvpMapPoints[int(best[q])] = qmp[q]
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
