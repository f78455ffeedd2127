"""Deterministic synthetic inputs.  numpy PCG64 with fixed seeds."""
import numpy as np


def rectangles_scene(nrows, ncols, seed=42, nrect=None, noise=4):
    """Axis-aligned rectangles of random size / grey level + uniform noise: corner density ~0.1-1 %."""
    rng = np.random.default_rng(seed)
    img = np.full((nrows, ncols), 128, dtype=np.int32)
    if nrect is None:
        nrect = max(8, (nrows * ncols) // 4000)
    for _ in range(nrect):
        h, w = rng.integers(8, 200, 2)
        r, c = rng.integers(0, nrows), rng.integers(0, ncols)
        img[r:r + h, c:c + w] = rng.integers(0, 256)
    img = img + rng.integers(-noise, noise + 1, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def _blur1461(a):
    k = np.array([1, 4, 6, 4, 1], dtype=np.float64) / 16
    a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, 2, mode="reflect"), k, mode="valid"), 0, a)
    a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, 2, mode="reflect"), k, mode="valid"), 1, a)
    return a


def smooth_noise(nrows, ncols, seed=42, sigmas=(1.5, 5.0, 14.0)):
    """Multi-octave band-limited noise in [0,1]: white noise Gaussian-blurred at several scales so
    that every pyramid level (factor 2, up to 4 levels) still carries texture."""
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(seed)
    a = np.zeros((nrows, ncols))
    for s in sigmas:
        o = gaussian_filter(rng.standard_normal((nrows, ncols)), sigma=s, mode="reflect")
        a += o / o.std()
    lo, hi = np.percentile(a, 0.5), np.percentile(a, 99.5)
    return np.clip((a - lo) / (hi - lo), 0, 1)


def warp(a, dr, dc):
    """Resample float image a at (r + dr, c + dc) with bilinear interpolation in float64 (generator only)."""
    nr, nc = a.shape
    rr, cc = np.meshgrid(np.arange(nr, dtype=np.float64), np.arange(nc, dtype=np.float64), indexing="ij")
    rr = np.clip(rr + dr, 0, nr - 1.001)
    cc = np.clip(cc + dc, 0, nc - 1.001)
    r0, c0 = rr.astype(np.int64), cc.astype(np.int64)
    a0, a1 = rr - r0, cc - c0
    return ((1 - a0) * (1 - a1) * a[r0, c0] + a0 * (1 - a1) * a[r0 + 1, c0] + (1 - a0) * a1 * a[r0, c0 + 1] + a0 * a1 * a[r0 + 1, c0 + 1])


def lk_pair(nrows, ncols, nkps, seed=42, shift=(2.3, -1.7), margin=24):
    """frame1 = band-limited noise; frame2(p) = frame1(p - flow(p)) with a smooth sub-pixel flow
    (global shift + 0.5 px sinusoid); keypoints on a jittered grid >= margin px from the edges."""
    rng = np.random.default_rng(seed)
    a = smooth_noise(nrows, ncols, seed)
    rr, cc = np.meshgrid(np.arange(nrows, dtype=np.float64), np.arange(ncols, dtype=np.float64), indexing="ij")
    fr = shift[0] + 0.5 * np.sin(rr / 97.0 + cc / 131.0)
    fc = shift[1] + 0.5 * np.cos(rr / 113.0 - cc / 89.0)
    b = warp(a, -fr, -fc)
    f1 = np.clip(np.round(a * 255), 0, 255).astype(np.uint8)
    f2 = np.clip(np.round(b * 255), 0, 255).astype(np.uint8)
    side = int(np.ceil(np.sqrt(nkps)))
    gr = np.linspace(margin, nrows - 1 - margin, side)
    gc = np.linspace(margin, ncols - 1 - margin, side)
    pts = np.stack(np.meshgrid(gr, gc, indexing="ij"), -1).reshape(-1, 2)[:nkps]
    pts = pts + rng.uniform(-3, 3, pts.shape)
    pts = np.floor(pts)  # the reference feeds integer-valued keypoints (vint2 / FAST output)
    pts[:, 0] = np.clip(pts[:, 0], margin, nrows - 2 - margin)
    pts[:, 1] = np.clip(pts[:, 1], margin, ncols - 2 - margin)
    return f1, f2, pts.astype(np.float32)
