"""Deterministic synthetic KITTI-shaped stereo pairs (KITTI itself is not available offline).

Left image = multi-octave box-filtered noise (flat, low-contrast top third so FAST cells there take the
iniThFAST -> minThFAST retry, reference ORBextractor.cpp:808-815) plus textured rectangles; every
rectangle and the background carry their own disparity, the right image is the same scene painted at
x - d with linear interpolation for the fractional part of d.  numpy only (no cv2) so the bench does
not depend on anything the product does not need.
"""
import hashlib

import numpy as np


def _box(a, r):
    """(2r+1)^2 box filter with edge replication, via cumulative sums."""
    if r <= 0:
        return a
    p = np.pad(a, r + 1, mode="edge")[:, :]
    c = np.cumsum(np.cumsum(p, axis=0, dtype=np.float64), axis=1)
    k = 2 * r + 1
    s = c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]
    return (s / (k * k))[: a.shape[0], : a.shape[1]]


def _texture(rng, h, w, speckle_every=380):
    """Smooth low-frequency texture plus sparse high-contrast speckles (the repeatable FAST corners)."""
    t = 0.35 * _box(rng.random((h, w)), 3) + 0.65 * _box(rng.random((h, w)), 9)
    t -= t.min()
    t = t / max(t.max(), 1e-9)
    n = max(1, (h * w) // speckle_every)
    ys = rng.integers(0, h, n)
    xs = rng.integers(0, w, n)
    val = rng.random(n)
    for y, x, v in zip(ys, xs, val):   # peaked 3x3 blobs: one FAST corner each, response ~ |v - local gray|
        t[y:y + 3, x:x + 3] = 0.5 * (t[y:y + 3, x:x + 3] + v)
        t[min(y + 1, h - 1), min(x + 1, w - 1)] = v
    return t


def _shift_cols(img, d):
    """img sampled at x + d (d >= 0 float): out[:, x] = (1-f) img[:, x+n] + f img[:, x+n+1], edge clamped."""
    n = int(np.floor(d))
    f = float(d - n)
    w = img.shape[1]
    xs = np.arange(w)
    a = img[:, np.clip(xs + n, 0, w - 1)]
    b = img[:, np.clip(xs + n + 1, 0, w - 1)]
    return (1.0 - f) * a + f * b


def make_stereo_pair(idx, H=376, W=1241, n_rect=60, max_disp=96.0):
    """Returns (left, right) uint8 [H, W] for frame `idx` (same idx -> same bytes)."""
    rng = np.random.default_rng(1000 + int(idx))
    margin = int(max_disp) + 8
    WW = W + margin                      # paint on a wider canvas so the right view has content at its edge
    bg = _texture(rng, H, WW)
    contrast = np.ones((H, 1))
    top = H // 3
    contrast[:top] = np.linspace(0.10, 0.35, top)[:, None]   # sky-like: low contrast
    canvas = 0.5 + (bg - 0.5) * contrast
    d_bg = float(rng.integers(2, 7))
    left = canvas.copy()
    right = _shift_cols(canvas, d_bg)
    # rectangles from far to near (disparity ascending) so nearer ones occlude farther ones
    ds = rng.uniform(6.0, max_disp, n_rect)
    ds = np.sort(np.where(rng.random(n_rect) < 0.6, np.floor(ds), ds))   # 60 % integer disparities
    for d in ds:
        rh = int(rng.integers(20, 110))
        rw = int(rng.integers(30, 220))
        y0 = int(rng.integers(top // 2, H - 4))
        x0 = int(rng.integers(0, WW - 4))
        y1, x1 = min(y0 + rh, H), min(x0 + rw, WW)
        g = float(rng.uniform(0.05, 0.95))
        tex = np.clip(g + 0.6 * (_texture(rng, y1 - y0, x1 - x0) - 0.5), 0, 1)
        left[y0:y1, x0:x1] = tex
        # right view: the rectangle sits at x0 - d; paint via a shifted copy of a sparse layer
        n = int(np.floor(d))
        f = float(d - n)
        xa0, xa1 = x0 - n - 1, x1 - n        # covers both integer taps
        if xa1 <= 0:
            continue
        layer = np.zeros((y1 - y0, x1 - x0 + 1))
        alpha = np.zeros((y1 - y0, x1 - x0 + 1))
        layer[:, 1:] += (1 - f) * tex
        alpha[:, 1:] += (1 - f)
        layer[:, :-1] += f * tex
        alpha[:, :-1] += f
        cx0 = max(xa0, 0)
        sub = slice(cx0 - xa0, xa1 - xa0)
        dst = right[y0:y1, cx0:xa1]
        dst[...] = dst * (1 - alpha[:, sub]) + layer[:, sub]
    left = left[:, :W]
    right = right[:, :W]
    to8 = lambda a: np.ascontiguousarray(np.clip(np.rint(a * 255.0), 0, 255).astype(np.uint8))
    return to8(left), to8(right)


def pair_digest(left, right):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(left).tobytes())
    h.update(np.ascontiguousarray(right).tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------
# KITTI-like scenes (the bench's default workload).  make_stereo_pair above is kept unchanged: the committed golden vectors
# and most parity tests are built on it, and it stays a second, harsher source (more FAST candidates, heavy occlusion).
# The statistics that matter for the kernels were tuned against the one real KITTI frame the reference ships
# (pyORBExtractor/kitti06-436.png: ~9.4 k FAST candidates per image, ~40 % of the 30-px cells need the minThFAST retry,
# ORBextractor.cpp:808-815) and against the match rate the reference's stereo matcher reaches on KITTI (~1400 of 2000):
#   * a smooth bright sky band, a low-contrast road plane with lane markings, textured structures in between;
#   * disparity: 2-4 px for the sky / far background, growing linearly with the row on the road plane, constant on every
#     upright object (= the road's disparity at the object's foot), objects sorted far to near;
#   * both views are composited layer by layer (right view = every layer shifted by its own disparity, sub-pixel part by
#     linear interpolation), then get independent sensor noise.
# ------------------------------------------------------------------------------------------------
def _shift_rows(img, d_rows):
    """img sampled at x + d_rows[y] per row (d >= 0 float array [H]), linear interpolation, edge clamped."""
    h, w = img.shape
    n = np.floor(d_rows).astype(np.int64)
    f = (d_rows - n)[:, None]
    xs = np.arange(w)[None, :]
    ia = np.clip(xs + n[:, None], 0, w - 1)
    ib = np.clip(xs + n[:, None] + 1, 0, w - 1)
    return (1.0 - f) * np.take_along_axis(img, ia, 1) + f * np.take_along_axis(img, ib, 1)


def _structure_texture(rng, h, w, contrast, kind=None):
    """Facade-like (smooth wall + rows of window blocks) or vegetation-like (fine mid-contrast texture) surface, zero mean."""
    kind = (2 if rng.random() < 0.15 else 0) if kind is None else kind
    if kind == 2:                                    # vegetation / gravel: fine texture, many weak corners
        t = 0.6 * _box(rng.random((h, w)), 1) + 0.4 * _box(rng.random((h, w)), 3)
        t = _box(t, 1)
        t = (t - t.mean()) / max(t.std(), 1e-9) * 0.09
        return contrast * t
    t = 0.5 * _box(rng.random((h, w)), 6) + 0.5 * _box(rng.random((h, w)), 18)
    t = (t - t.mean()) / max(t.std(), 1e-9) * 0.035          # smooth wall shading
    # windows / panels: constant blocks of individual position, size and contrast (no periodic grid: a regular facade makes
    # the row-band Hamming search pick the neighbouring window and the SAD refinement reject it)
    for _ in range(max(1, (h * w) // 700)):
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        bh, bw_ = int(rng.integers(4, 22)), int(rng.integers(4, 26))
        t[y:y + bh, x:x + bw_] = t[y, x] + rng.normal(0.0, 0.2)
    # individual details: signs, pipes, shadows
    for _ in range(max(1, (h * w) // 1700)):
        y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
        bh, bw_ = int(rng.integers(2, 25)), int(rng.integers(2, 25))
        t[y:y + bh, x:x + bw_] += rng.normal(0.0, 0.16)
    return contrast * _box(t, 1) * 1.5


def make_kitti_like_pair(idx, H=376, W=1241, max_disp=88.0, noise_sigma=0.6, nobj_range=(10, 16), obj_scale=1.3):
    """Returns (left, right) uint8 [H, W] for frame `idx`: a road scene with KITTI-like corner / retry / match statistics."""
    rng = np.random.default_rng(77000 + int(idx))
    margin = int(max_disp) + 8
    WW = W + margin
    yh = int(H * rng.uniform(0.42, 0.50))                       # horizon row
    g = max_disp / max(H - 1 - yh, 1)                           # road-plane disparity per row below the horizon
    rows = np.arange(H)
    # ---- layer 0: sky + far background band (disparity 2-4 px) ----
    sky = 0.86 - 0.25 * (rows / max(yh, 1))[:, None] + 0.012 * (_box(rng.random((H, WW)), 10) - 0.5) * 6
    far_h = int(rng.integers(18, 40))
    far = 0.42 + _structure_texture(rng, H, WW, 0.8, kind=2)
    skyline = yh - far_h + (8 * (_box(rng.random((1, WW)), 30)[0] - 0.5) * 6).astype(int)
    layer0 = np.where(rows[:, None] >= skyline[None, :], far, sky)
    d0v = float(rng.uniform(2.0, 4.0))                             # bf = 386 px m: 100-200 m
    d0 = np.full(H, d0v)
    left = layer0.copy()
    right = _shift_rows(layer0, d0)
    # ---- layer 1: road plane below the horizon, disparity g * (y - yh) ----
    road = 0.34 + 0.06 * (_box(rng.random((H, WW)), 1) - 0.5) * 2 + 0.05 * (_box(rng.random((H, WW)), 6) - 0.5) * 4
    cx = WW * rng.uniform(0.42, 0.58)
    for k, off in enumerate((-1.6, -0.05, 0.05, 1.6)):           # lane markings converge at the vanishing point
        for y in range(yh + 4, H):
            s = (y - yh) / max(H - yh, 1)
            xc = cx + off * s * W * 0.42
            hw = 0.5 + 3.5 * s
            if k in (1, 2) and ((y - yh) // max(int(6 + 30 * s), 1)) % 2:
                continue                                         # dashed centre line
            x0, x1 = int(xc - hw), int(xc + hw) + 1
            if x1 > 0 and x0 < WW:
                road[y, max(x0, 0):min(x1, WW)] = 0.78
    d1 = np.maximum(g * (rows - yh), d0v)
    m1 = (rows >= yh)[:, None] & np.ones((1, WW), bool)
    left = np.where(m1, road, left)
    right = np.where(m1, _shift_rows(road, d1), right)
    # ---- upright objects (buildings, trunks, cars, signs), far to near ----
    nobj = int(rng.integers(*nobj_range))
    feet = np.sort(rng.uniform(yh + 3, yh + 0.62 * (H - yh), nobj))
    for yb in feet:
        d = max(float(g * (yb - yh)), d0v)
        scale = (yb - yh) / max(H - yh, 1)
        oh = int(rng.uniform(0.35, 1.0) * (40 + 420 * scale) * obj_scale)
        ow = int(rng.uniform(0.3, 1.0) * (60 + 520 * scale) * obj_scale)
        side = rng.random() < 0.8                                 # most objects flank the road
        if side:
            x0 = int(rng.uniform(0, max(0.30 * WW - ow * 0.5, 1.0))) if rng.random() < 0.5 else int(rng.uniform(min(0.70 * WW - ow * 0.5, WW - 9.0), WW - 8))
        else:
            x0 = int(rng.uniform(0.30 * WW, max(0.70 * WW - 4, 0.30 * WW + 1)))
        y1, y0 = int(yb), max(int(yb) - oh, 0)
        x0 = max(x0, 0)
        x1 = min(x0 + max(ow, 8), WW)
        if y1 - y0 < 4 or x1 - x0 < 4:
            continue
        base_gray = float(rng.uniform(0.15, 0.7))
        tex = np.clip(base_gray + _structure_texture(rng, y1 - y0, x1 - x0, float(rng.uniform(0.6, 1.0))), 0, 1)
        left[y0:y1, x0:x1] = tex
        n = int(np.floor(d))
        f = float(d - n)
        xa0, xa1 = x0 - n - 1, x1 - n
        if xa1 <= 0:
            continue
        layer = np.zeros((y1 - y0, x1 - x0 + 1))
        alpha = np.zeros((y1 - y0, x1 - x0 + 1))
        layer[:, 1:] += (1 - f) * tex
        alpha[:, 1:] += (1 - f)
        layer[:, :-1] += f * tex
        alpha[:, :-1] += f
        c0 = max(xa0, 0)
        sub = slice(c0 - xa0, xa1 - xa0)
        dst = right[y0:y1, c0:xa1]
        dst[...] = dst * (1 - alpha[:, sub]) + layer[:, sub]
    left, right = left[:, :W], right[:, :W]
    # slight optical blur + independent sensor noise per view
    nl = rng.normal(0.0, noise_sigma, (H, W))
    nr = rng.normal(0.0, noise_sigma, (H, W))
    to8 = lambda a, nz: np.ascontiguousarray(np.clip(np.rint(a * 255.0 + nz), 0, 255).astype(np.uint8))
    return to8(left, nl), to8(right, nr)
