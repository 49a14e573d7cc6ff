"""Synthetic stereo inputs of the benchmark shapes.

Deterministic per frame index f (`np.random.default_rng(1000 + f)`): blurred-noise images whose right view is the left
view shifted by a smooth ground-truth disparity, sparse hints at 5 % density (uniform random, or LiDAR-like scan
lines), and the VPP parameters the configs name.  No dataset or network access is needed.
"""
import numpy as np

SHAPES = {"V": (480, 640), "K": (375, 1242), "M": (1988, 2880)}


def _blur5(img, sigma=1.2):
    """Separable 5-tap Gaussian (reflect-101 border), uint8 in/out with round-half-up: texture only needs to be
    deterministic, not equal to any library's blur."""
    k = np.exp(-0.5 * (np.arange(-2, 3) / sigma) ** 2)
    k /= k.sum()
    a = img.astype(np.float32)
    for axis in (0, 1):
        p = np.pad(a, [(2, 2) if ax == axis else (0, 0) for ax in range(a.ndim)], mode="reflect")
        a = sum(k[i] * np.take(p, range(i, i + a.shape[axis]), axis=axis) for i in range(5))
    return np.clip(np.floor(a + 0.5), 0, 255).astype(np.uint8)


def make_pair(f, shape="K", hints="lidar", channels=3, density=0.05, foreground=0):
    """Returns dict(left, right uint8 [H,W,C]; gt float32 [H,W]; hints float32 [H,W] (0 = none)).
    foreground = k adds k fronto-parallel boxes 20..45 px closer than the background, so that background hints next to
    their left edges are occluded in the right view (the case filter.occlusion_heuristic exists for)."""
    H, W = SHAPES[shape] if isinstance(shape, str) else shape
    rng = np.random.default_rng(1000 + f)
    base = _blur5(rng.integers(0, 256, (H, W + 256, channels), dtype=np.uint8))
    left = np.ascontiguousarray(base[:, 128:128 + W])
    yy, xx = np.mgrid[0:H, 0:W]
    dgt = (8.0 + 120.0 * yy / max(H - 1, 1) + 6.0 * np.sin(xx / 97.0)).astype(np.float32)
    if foreground:
        brng = np.random.default_rng(77000 + f)
        for _ in range(int(foreground)):
            bh, bw = int(brng.integers(max(H // 8, 4), max(H // 3, 6))), int(brng.integers(max(W // 12, 4), max(W // 4, 6)))
            by, bx = int(brng.integers(0, max(H - bh, 1))), int(brng.integers(0, max(W - bw, 1)))
            dgt[by:by + bh, bx:bx + bw] = np.float32(dgt[by:by + bh, bx:bx + bw].max() + brng.uniform(20, 45))
        dgt = np.minimum(dgt, np.float32(190.0))
    src = np.clip(128 + xx + np.rint(dgt).astype(np.int64), 0, W + 255)
    right = np.ascontiguousarray(base[yy, src])
    if hints == "random":
        mask = rng.random((H, W)) < density
    else:  # LiDAR-like: every 4th row from y=120 (scaled for small frames), 30 % of the columns
        y0 = min(120, H // 3)
        rows = (yy >= y0) & ((yy - y0) % 4 == 0)
        mask = rows & (rng.random((H, W)) < 0.30)
    g = np.where(mask, dgt + rng.normal(0, 0.25, (H, W)), 0).astype(np.float32)
    g = np.where(mask, np.maximum(g, np.float32(0.5)), 0).astype(np.float32)
    if channels == 1:
        left, right = left[..., 0], right[..., 0]
    return dict(left=left, right=right, gt=dgt, hints=g)


def make_batch(n, f0=0, **kw):
    frames = [make_pair(f0 + i, **kw) for i in range(n)]
    return {k: np.stack([fr[k] for fr in frames]) for k in frames[0]}
