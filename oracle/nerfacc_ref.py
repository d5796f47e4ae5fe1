"""oracle/nerfacc_ref.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Restatement of the slice of third-party ``nerfacc`` 0.3.x that the reference calls
(``nerf/run_nerf_acc.py:12,197-198``, ``nerf/nerf_helpers_acc.py:29,72-76``,
``visualization/visualization.py:162,214``).  nerfacc is not in /root/reference, not
installable here, CUDA-only and unpinned by the reference => PARITY UNPINNED; this file
(and oracle/march_ref.c, which holds the per-ray arithmetic) is the canonical definition
the CUDA kernels are compared against.  Python-level composition follows nerfacc 0.3.5's
``ray_marching.py`` / ``grid.py`` / ``vol_rendering.py``:

  ray_marching:  slab test -> clamp to [near_plane, far_plane] -> two-pass march through the
                 binary grid -> alpha_fn on all samples -> render_visibility with
                 alpha_thre := min(alpha_thre, mean(grid.occs)) -> boolean-mask compaction.
  OccupancyGrid: occs fp32 [res^3] (init 0), binary bool [res,res,res] (init False);
                 every_n_step(step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95,
                 warmup_steps=256, n=16).
"""
import ctypes
import numpy as np
import torch

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class OccupancyGrid:
    """nerfacc 0.3.x OccupancyGrid (ContractionType.AABB only), numpy state."""

    def __init__(self, roi_aabb, resolution=128):
        self.roi_aabb = _f32(roi_aabb).reshape(6)
        self.resolution = int(resolution)
        self.num_cells = self.resolution ** 3
        self.occs = np.zeros(self.num_cells, dtype=np.float32)
        self.binary = np.zeros((self.resolution,) * 3, dtype=bool)
        r = np.arange(self.resolution)
        self.grid_coords = np.stack(np.meshgrid(r, r, r, indexing="ij"), axis=-1).reshape(-1, 3)

    # -- nerfacc grid.py: _sample_uniform_and_occupied_cells -------------------------------
    def sample_cells(self, step, rng, warmup_steps=256):
        if step < warmup_steps:
            return np.arange(self.num_cells, dtype=np.int64)
        n = self.num_cells // 4
        uniform = rng.integers(0, self.num_cells, size=n, dtype=np.int64)
        occupied = np.nonzero(self.binary.reshape(-1))[0].astype(np.int64)
        if n < len(occupied):
            occupied = occupied[rng.integers(0, len(occupied), size=n)]
        return np.concatenate([uniform, occupied])

    def cell_points(self, indices, jitter):
        """x = (grid_coords + U[0,1)^3) / res mapped into the AABB (contract_inv)."""
        lo, hi = self.roi_aabb[:3], self.roi_aabb[3:]
        x = (self.grid_coords[indices].astype(np.float32) + _f32(jitter)) / np.float32(self.resolution)
        return (x * (hi - lo) + lo).astype(np.float32)

    def update_from_occ(self, indices, occ, occ_thre=0.01, ema_decay=0.95):
        """occs[idx] = max(occs[idx]*decay, occ); binary = occs > min(mean(occs), occ_thre).

        With duplicate indices numpy keeps the LAST write (the library's GPU scatter is
        non-deterministic there; the CUDA kernel takes the max over duplicates instead --
        tests use duplicate-free indices for bit-exact comparison).
        """
        occ = _f32(occ).reshape(-1)
        self.occs[indices] = np.maximum(self.occs[indices] * np.float32(ema_decay), occ)
        thre = min(float(self.occs.mean(dtype=np.float32)), float(np.float32(occ_thre)))
        self.binary = (self.occs > np.float32(thre)).reshape(self.binary.shape)

    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16,
                     rng=None, indices=None, jitter=None):
        if step % n != 0:
            return
        rng = rng if rng is not None else np.random.default_rng(0)
        if indices is None:
            indices = self.sample_cells(step, rng, warmup_steps)
        if jitter is None:
            jitter = rng.random((len(indices), 3), dtype=np.float32)
        x = self.cell_points(indices, jitter)
        occ = occ_eval_fn(x)
        self.update_from_occ(indices, np.asarray(occ).reshape(-1), occ_thre, ema_decay)

    def query_occ(self, pts):
        pts = _f32(pts).reshape(-1, 3)
        out = np.empty(len(pts), dtype=np.float32)
        b = np.ascontiguousarray(self.binary.astype(np.uint8))
        lib().oracle_grid_query(ctypes.c_int64(len(pts)), _p(pts), _p(self.roi_aabb),
                                ctypes.c_int(self.resolution), _p(b), _p(out))
        return out


def ray_aabb_intersect(rays_o, rays_d, aabb, near_plane, far_plane):
    o, d, aabb = _f32(rays_o), _f32(rays_d), _f32(aabb)
    n = len(o)
    t_min = np.empty(n, np.float32)
    t_max = np.empty(n, np.float32)
    lib().oracle_ray_aabb_intersect(ctypes.c_int64(n), _p(o), _p(d), _p(aabb), ctypes.c_float(near_plane),
                                    ctypes.c_float(far_plane), _p(t_min), _p(t_max))
    return t_min, t_max


def march(rays_o, rays_d, t_min, t_max, roi, resolution, binary, step_size):
    """Two-pass march.  Returns (ray_indices[n] i64, t_starts[n] f32, t_ends[n] f32, offsets[R+1] i64)."""
    o, d = _f32(rays_o), _f32(rays_d)
    t_min, t_max, roi = _f32(t_min), _f32(t_max), _f32(roi)
    b = np.ascontiguousarray(np.asarray(binary).astype(np.uint8))
    n = len(o)
    counts = np.empty(n, np.int32)
    L = lib()
    L.oracle_march_count(ctypes.c_int64(n), _p(o), _p(d), _p(t_min), _p(t_max), _p(roi), ctypes.c_int(resolution),
                         _p(b), ctypes.c_float(step_size), _p(counts))
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    total = int(offsets[-1])
    ray_indices = np.empty(total, np.int64)
    t_starts = np.empty(total, np.float32)
    t_ends = np.empty(total, np.float32)
    L.oracle_march_write(ctypes.c_int64(n), _p(o), _p(d), _p(t_min), _p(t_max), _p(roi), ctypes.c_int(resolution),
                         _p(b), ctypes.c_float(step_size), _p(offsets), _p(ray_indices), _p(t_starts), _p(t_ends))
    return ray_indices, t_starts, t_ends, offsets


def visibility(offsets, alphas, early_stop_eps, alpha_thre):
    alphas = _f32(alphas).reshape(-1)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    keep = np.empty(len(alphas), np.uint8)
    lib().oracle_visibility(ctypes.c_int64(len(offsets) - 1), _p(offsets), _p(alphas),
                            ctypes.c_float(early_stop_eps), ctypes.c_float(alpha_thre), _p(keep))
    return keep.astype(bool)


def scatter_mul(src, index, n_rays):
    src = _f32(src).reshape(-1)
    index = np.ascontiguousarray(index, dtype=np.int64)
    out = np.empty(n_rays, np.float32)
    lib().oracle_scatter_mul(ctypes.c_int64(len(src)), _p(src), _p(index), ctypes.c_int64(n_rays), _p(out))
    return out


def ray_marching(rays_o, rays_d, scene_aabb, grid, alpha_fn, near_plane, far_plane,
                 early_stop_eps=1e-4, alpha_thre=0.0, render_step_size=1e-3, return_prefilter=False):
    """nerfacc.ray_marching(...) as called at /root/reference/nerf/nerf_helpers_acc.py:29.

    Returns (ray_indices[n'] int64, t_starts[n',1], t_ends[n',1]) as numpy arrays.
    """
    t_min, t_max = ray_aabb_intersect(rays_o, rays_d, scene_aabb, near_plane, far_plane)
    ray_indices, t_starts, t_ends, offsets = march(rays_o, rays_d, t_min, t_max, grid.roi_aabb, grid.resolution,
                                                   grid.binary, np.float32(render_step_size))
    pre = (ray_indices, t_starts[:, None], t_ends[:, None], offsets)
    if (alpha_thre > 0.0 or early_stop_eps > 0.0) and alpha_fn is not None and len(ray_indices) > 0:
        alphas = alpha_fn(t_starts[:, None], t_ends[:, None], ray_indices)
        thre = min(float(alpha_thre), float(grid.occs.mean(dtype=np.float32)))
        keep = visibility(offsets, np.asarray(alphas), early_stop_eps, thre)
        ray_indices, t_starts, t_ends = ray_indices[keep], t_starts[keep], t_ends[keep]
    out = (ray_indices, t_starts[:, None], t_ends[:, None])
    return out + (pre,) if return_prefilter else out
