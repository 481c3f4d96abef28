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
