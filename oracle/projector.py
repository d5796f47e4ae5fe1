"""oracle/projector.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference's ground-truth projector restated in numpy float64: ray_tracing (/root/reference/phantomdata/helpers.py:192-224)
on top of a scipy RegularGridInterpolator(bounds_error=False, fill_value=0) -- the interpolator the phantom scripts build
(/root/reference/phantomdata/cttoray.py).  tests/test_projector.py pins `trilinear` against scipy itself.
"""
import numpy as np


def trilinear(volume, lo, hi, pts):
    """Trilinear interpolation of volume[X,Y,Z] spanning [lo, hi] at pts[N,3]; 0 outside (fill_value=0)."""
    volume = np.asarray(volume, dtype=np.float64)
    n = np.array(volume.shape)
    g = (np.asarray(pts, dtype=np.float64) - lo) / (np.asarray(hi, dtype=np.float64) - lo) * (n - 1)
    inside = np.all((g >= 0) & (g <= n - 1), axis=1)
    i0 = np.clip(np.floor(g).astype(np.int64), 0, n - 2)
    f = g - i0
    out = np.zeros(len(g))
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                out += w * volume[i0[:, 0] + dx, i0[:, 1] + dy, i0[:, 2] + dz]
    return np.where(inside, out, 0.0)


def ray_tracing(volume, lo, hi, rays_o, rays_d, depths, kind="ct"):
    """helpers.py:192-224 for a flat list of rays: product over depths of exp(-mu * dist * |d|) ('ct') or exp(-mu) ('sdf')."""
    rays_o = np.asarray(rays_o, dtype=np.float64); rays_d = np.asarray(rays_d, dtype=np.float64)
    depths = np.asarray(depths, dtype=np.float64)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * depths[None, :, None]
    mu = trilinear(volume, np.asarray(lo, np.float64), np.asarray(hi, np.float64), pts.reshape(-1, 3)).reshape(len(rays_o), len(depths))
    if kind == "ct":
        dists = np.concatenate([depths[1:] - depths[:-1], [1e10]])                  # helpers.py:202
        w = np.exp(-mu * dists[None, :] * np.linalg.norm(rays_d, axis=1)[:, None])   # :208-211
    else:
        w = np.exp(-mu)                                                             # :213-215
    return np.prod(w, axis=1)
