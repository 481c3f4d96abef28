"""TEST INFRASTRUCTURE -- Python restatement of /root/reference/Frame.py:161-279 (compute_stereo_matches)
and :324-326 (descriptor_distance), keeping the reference's per-keypoint Python/NumPy structure so that
its run time is representative of the reference CPU path (the reference itself cannot travel to the GPU
box).  Numerics follow NumPy >= 2 scalar promotion: `mb`, `maxD`, `deltaR`,
`bestuR`, disparity and depth are float32.  Pinned against fixtures produced by the reference's own method
(tests/golden/make_golden.py); the C twin is `orbo_stereo` in orb_oracle.cpp.
"""
import math

import numpy as np

TH_HIGH = 100   # ORBMatcher.py:3
TH_LOW = 50     # ORBMatcher.py:4


def hamming_bytes(a, b):
    """Frame.py:324-326: popcount of the XOR, byte by byte through Python's bin()."""
    return sum(bin(v).count("1") for v in np.bitwise_xor(a, b))


def stereo_matches(keysL, descL, keysR, descR, scale_factors, inv_scale_factors, pyrL, pyrR, mbf, fx32):
    """keys*: sequences of (x, y, octave) with python-float x, y (widened float32) and int octave.
    pyr*: the GetImagePyramid() views.  Returns (mvuRight, mvDepth) as lists (-1 where unmatched)."""
    n_left = len(keysL)
    u_right = [-1] * n_left
    depth = [-1] * n_left
    th_orb = (TH_HIGH + TH_LOW) / 2
    n_rows = pyrL[0].shape[0]

    # Frame.py:170-179 -- which right keypoints may match a given image row
    row_lists = [[] for _ in range(n_rows)]
    for iR, (xr, yr, oR) in enumerate(keysR):
        reach = 2.0 * scale_factors[oR]
        for row in range(math.floor(yr - reach), math.ceil(yr + reach) + 1):
            row_lists[row].append(iR)

    mb = mbf / fx32            # python float / np.float32 -> np.float32 (Frame.py:43)
    min_d = 0
    max_d = mbf / mb           # Frame.py:181-183

    for iL, (uL, vL, oL) in enumerate(keysL):
        candidates = row_lists[int(vL)]
        if not candidates:
            continue
        lo_u = uL - max_d
        hi_u = uL - min_d
        if hi_u < 0:
            continue
        best, best_r = TH_HIGH, 0
        dl = descL[iL]
        for iR in candidates:
            xr, _, oR = keysR[iR]
            if oR < oL - 1 or oR > oL + 1:
                continue
            if lo_u <= xr <= hi_u:
                d = hamming_bytes(dl, descR[iR])
                if d < best:
                    best, best_r = d, iR
        if not best < th_orb:
            continue

        # Frame.py:224-269 -- 11x11 SAD slide on the keypoint's pyramid level
        inv = inv_scale_factors[oL]
        su = round(uL * inv)
        sv = round(vL * inv)
        sr = round(keysR[best_r][0] * inv)
        half, reach = 5, 5
        imL, imR = pyrL[oL], pyrR[oL]
        patch = imL[sv - half:sv + half + 1, su - half:su + half + 1].astype(np.float32)
        patch = patch - patch[half, half] * np.ones_like(patch, dtype=np.float32)
        if sr + reach - half < 0 or sr + reach + half + 1 >= imR.shape[1]:
            continue
        sad = [0] * (2 * reach + 1)
        best_sad, best_inc = float("inf"), 0
        for inc in range(-reach, reach + 1):
            other = imR[sv - half:sv + half + 1, sr + inc - half:sr + inc + half + 1].astype(np.float32)
            other = other - other[half, half] * np.ones_like(other, dtype=np.float32)
            s = np.sum(np.abs(patch - other))
            if s < best_sad:
                best_sad, best_inc = s, inc
            sad[reach + inc] = s
        if best_inc == -reach or best_inc == reach:
            continue
        d1, d2, d3 = sad[reach + best_inc - 1], sad[reach + best_inc], sad[reach + best_inc + 1]
        delta = (d1 - d3) / (2.0 * (d1 + d3 - 2.0 * d2))
        if delta < -1 or delta > 1:
            continue
        best_u = scale_factors[oL] * (sr + best_inc + delta)
        disparity = uL - best_u
        if min_d <= disparity < max_d:
            if disparity <= 0:
                disparity = 0.01
                best_u = uL - 0.01
            depth[iL] = mbf / disparity
            u_right[iL] = best_u
    return u_right, depth
