"""Validated Python entry points over the C ABI (one function per exported kernel group).

Tensors are allocated by torch and passed as raw device pointers; every call is enqueued on the current
torch CUDA stream.  Shape / dtype / device / contiguity checks happen here, before the C call.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import MlpDesc, Samples, OUT_ALPHA, OUT_LOGIT, OUT_SIGMA, PREC_BF16, PREC_FP32  # noqa: F401


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype, name, ndim=None, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError(f"{name} is required")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")
    return t


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _host6(v, name):
    a = np.ascontiguousarray(np.asarray(v.detach().cpu() if isinstance(v, torch.Tensor) else v, dtype=np.float32)).reshape(-1)
    if a.size != 6:
        raise ValueError(f"{name} must have 6 elements")
    return a


def _alloc(pool, key, n, dtype, device):
    """n elements: a fresh tensor, or -- with a BufferPool -- a view of the pooled buffer `key` (valid until the next request
    for the same key; the sync-free training loop uses this so that a steady-state step never touches the CUDA allocator)."""
    if pool is None:
        return torch.empty((int(n),), dtype=dtype, device=device)
    return pool.typed(key, int(n), dtype, device)


def sm_count() -> int:
    return int(_lib.load().angio_sm_count())


# ------------------------------------------------------------------------------------------------ ray generation
def raygen(cam2world, img_w, img_h, focal, view=0, view_ids=None, px=None, py=None, pixels=None):
    """Cone-beam rays.  Image mode (view) -> (o[H*W,3], d[H*W,3]); gather mode (view_ids, px, py) -> (o, d[, pix])."""
    lib = _lib.load()
    cam2world = _chk(cam2world, torch.float64, "cam2world", 3)
    dev = cam2world.device
    if view_ids is None:
        n = int(img_w) * int(img_h)
    else:
        view_ids = _chk(view_ids, torch.int32, "view_ids", 1)
        px = _chk(px, torch.int32, "px", 1)
        py = _chk(py, torch.int32, "py", 1)
        n = view_ids.numel()
    pixels = _chk(pixels, torch.float32, "pixels", 3, allow_none=True)
    o = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d = torch.empty((n, 3), dtype=torch.float32, device=dev)
    pix = torch.empty((n,), dtype=torch.float32, device=dev) if (pixels is not None and view_ids is not None) else None
    _lib.check(lib.angio_raygen(_p(cam2world), int(view), _p(view_ids), _p(px), _p(py), n, int(img_w), int(img_h), float(focal),
                                _p(pixels) if pix is not None else None, _p(o), _p(d), _p(pix), _stream()), "angio_raygen")
    return (o, d, pix) if pix is not None else (o, d)


def sample_without_replacement(n, n_pool, weights, wsum, wsum2, seed, device, pool=None, status=None):
    """n ids out of n_pool, weighted, without replacement, uniformly shuffled (exponential race + threshold pre-filter + exact
    radix select + hash-bucket shuffle).  Two kernels, no host sync.  Returns (ids int64 [n], status int32 [2]): status[1] != 0
    means the candidate buffer overflowed / under-filled (astronomically unlikely by construction; callers that sync anyway check it).

    The expected number of rays with key < tau is sum_i (1 - exp(-w_i tau)), bracketed by tau*S1 - tau^2*S2/2 and tau*S1
    (S1 = sum w, S2 = sum w^2): tau is chosen from the lower bound so at least n + 8 sigma candidates are expected, the
    candidate buffers are sized from the upper bound.  Small pools / large sampling fractions consider every ray."""
    import math
    lib = _lib.load()
    weights = _chk(weights, torch.float32, "weights", 1, allow_none=True)
    target = n + 8.0 * math.sqrt(n) + 32.0
    disc = wsum * wsum - 2.0 * wsum2 * target
    if n > 0.1 * n_pool or disc <= 0.0:
        tau, capacity = 3.0e38, int(n_pool)
    else:
        tau = (wsum - math.sqrt(disc)) / wsum2
        capacity = min(int(n_pool), int(tau * wsum + 10.0 * math.sqrt(tau * wsum) + 64))
    capacity = max(capacity, int(n))
    wb = int(lib.angio_sample_rays_workspace_bytes(capacity, int(n)))
    ws = pool.get("sampler_ws", wb, device) if pool is not None else torch.empty((wb,), dtype=torch.uint8, device=device)
    ids = torch.empty((n,), dtype=torch.int64, device=device)
    status = torch.empty((2,), dtype=torch.int32, device=device) if status is None else _chk(status, torch.int32, "status", 1)
    _lib.check(lib.angio_sample_rays(_p(weights), int(n_pool), int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, float(tau), capacity, _p(ids),
                                     _p(status), _p(ws), wb, _stream()), "angio_sample_rays")
    return ids, status


def sample_select(keys, cand_ids, n, seed):
    """Check mode of the sampler: the n candidates with the smallest (positive fp32) keys, shuffled like a draw -- the selection
    half of `sample_without_replacement` on caller-supplied race keys.  Equal keys at the threshold are taken in id order.
    Returns (ids int64 [n], status int32 [2])."""
    lib = _lib.load()
    keys = _chk(keys, torch.float32, "keys", 1)
    cand_ids = _chk(cand_ids, torch.int64, "cand_ids", 1)
    m = keys.numel()
    if cand_ids.numel() != m:
        raise ValueError("keys and cand_ids must have the same length")
    wb = int(lib.angio_sample_rays_workspace_bytes(m, int(n)))
    ws = torch.empty((wb,), dtype=torch.uint8, device=keys.device)
    ids = torch.empty((n,), dtype=torch.int64, device=keys.device)
    status = torch.empty((2,), dtype=torch.int32, device=keys.device)
    _lib.check(lib.angio_sample_select(_p(keys), _p(cand_ids), m, int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(ids), _p(status), _p(ws), wb,
                                       _stream()), "angio_sample_select")
    return ids, status


def raygen_flat(cam2world, ids, img_w, img_h, focal, pixels=None):
    """Rays (o, d[, target pixel]) of flat ray ids (view * H * W + y * W + x): the gather that follows the sampler."""
    lib = _lib.load()
    cam2world = _chk(cam2world, torch.float64, "cam2world", 3)
    ids = _chk(ids, torch.int64, "ids", 1)
    pixels = _chk(pixels, torch.float32, "pixels", 3, allow_none=True)
    n, dev = ids.numel(), ids.device
    o = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d = torch.empty((n, 3), dtype=torch.float32, device=dev)
    pix = torch.empty((n,), dtype=torch.float32, device=dev) if pixels is not None else None
    _lib.check(lib.angio_raygen_flat(_p(cam2world), _p(ids), n, int(img_w), int(img_h), float(focal), _p(pixels), _p(o), _p(d), _p(pix),
                                     _stream()), "angio_raygen_flat")
    return (o, d, pix) if pix is not None else (o, d)


# ------------------------------------------------------------------------------------------------ marching
def exclusive_scan(counts, total_out=None):
    """offsets[n+1] (offsets[n] = total); total_out (optional int32[1] device tensor) also receives the total."""
    lib = _lib.load()
    counts = _chk(counts, torch.int32, "counts", 1)
    total_out = _chk(total_out, torch.int32, "total_out", 1, allow_none=True)
    n = counts.numel()
    offsets = torch.empty((n + 1,), dtype=torch.int32, device=counts.device)
    _lib.check(lib.angio_exclusive_scan_i32(_p(counts), n, _p(offsets), _p(total_out), _stream()), "angio_exclusive_scan_i32")
    return offsets


def march(rays_o, rays_d, scene_aabb, roi_aabb, resolution, binary, near_plane, far_plane, step_size, capacity=None, total_out=None,
          use_runs=True, pool=None, tag="march"):
    """Two-pass occupancy-grid march.  Returns (ray_idx int32 [n], t_starts [n], t_ends [n], offsets int32 [R+1]).

    capacity=None: one host sync (reads the total sample count to size the outputs), like the reference library.
    capacity=C   : NO host sync -- the sample arrays have C entries, the first offsets[R] of which are valid; C must be an
                   upper bound (R * (ceil((far - near) / step) + 1) always is).  Downstream kernels read the count on the
                   device (offsets[R], also written to total_out when given).
    use_runs: the count pass records each ray's runs of consecutive samples and the write pass replays them (no second grid
              march); False re-marches in the write pass.  Bit-identical samples either way.
    """
    lib = _lib.load()
    rays_o = _chk(rays_o, torch.float32, "ray_origins", 2)
    rays_d = _chk(rays_d, torch.float32, "ray_directions", 2)
    if rays_o.shape != rays_d.shape or rays_o.shape[1] != 3:
        raise ValueError("ray_origins / ray_directions must both be [R, 3]")
    binary = _bin_u8(binary, resolution)
    aabb = _host6(scene_aabb, "scene_aabb")
    roi = _host6(roi_aabb, "roi_aabb")
    R, dev = rays_o.shape[0], rays_o.device
    if R == 0:
        e = torch.empty((0,), dtype=torch.float32, device=dev)
        return torch.empty((0,), dtype=torch.int32, device=dev), e, e.clone(), torch.zeros((1,), dtype=torch.int32, device=dev)
    t_min = torch.empty((R,), dtype=torch.float32, device=dev)
    t_max = torch.empty((R,), dtype=torch.float32, device=dev)
    counts = torch.empty((R,), dtype=torch.int32, device=dev)
    runs = _alloc(pool, tag + "_runs", int(lib.angio_march_runs_bytes(R)), torch.uint8, dev) if use_runs else None
    _lib.check(lib.angio_march_count(_p(rays_o), _p(rays_d), R, aabb.ctypes.data, roi.ctypes.data, int(resolution), _p(binary),
                                     float(near_plane), float(far_plane), float(step_size), _p(t_min), _p(t_max), _p(counts),
                                     _p(runs), None, _stream()), "angio_march_count")
    offsets = exclusive_scan(counts, total_out)
    n = int(offsets[-1].item()) if capacity is None else int(capacity)
    ray_idx = _alloc(pool, tag + "_idx", n, torch.int32, dev)
    t0 = _alloc(pool, tag + "_t0", n, torch.float32, dev)
    t1 = _alloc(pool, tag + "_t1", n, torch.float32, dev)
    if n > 0:
        _lib.check(lib.angio_march_write(_p(rays_o), _p(rays_d), R, roi.ctypes.data, int(resolution), _p(binary), float(step_size),
                                         _p(t_min), _p(t_max), _p(offsets), _p(runs), n, _p(ray_idx), _p(t0), _p(t1), _stream()),
                   "angio_march_write")
    return ray_idx, t0, t1, offsets


def march_capacity(n_rays, near_plane, far_plane, step_size):
    """Upper bound of the samples `march` can emit for n_rays rays: every sample advances t by step_size inside [near, far]."""
    import math
    return int(n_rays) * (int(math.ceil((float(far_plane) - float(near_plane)) / float(step_size))) + 2)


def _bin_u8(binary, resolution):
    if not isinstance(binary, torch.Tensor) or not binary.is_cuda:
        raise ValueError("grid binary must be a CUDA tensor")
    if binary.dtype == torch.bool:
        binary = binary.contiguous().view(torch.uint8)
    binary = _chk(binary, torch.uint8, "binary")
    if binary.numel() != int(resolution) ** 3:
        raise ValueError("binary must have resolution^3 cells")
    return binary


def grid_query(points, roi_aabb, resolution, binary):
    lib = _lib.load()
    points = _chk(points, torch.float32, "points", 2)
    binary = _bin_u8(binary, resolution)
    roi = _host6(roi_aabb, "roi_aabb")
    out = torch.empty((points.shape[0],), dtype=torch.float32, device=points.device)
    _lib.check(lib.angio_grid_query(_p(points), points.shape[0], roi.ctypes.data, int(resolution), _p(binary), _p(out), _stream()),
               "angio_grid_query")
    return out


# ------------------------------------------------------------------------------------------------ visibility
def visibility_compact(alphas, offsets, t_starts, t_ends, early_stop_eps, alpha_thre, totals=None, capacity=None, pool=None, thre_cap=None):
    """Visibility mask + compaction.  Returns (ray_idx', t_starts', t_ends', offsets', keep).

    alphas / t_starts / t_ends may be capacity-sized (see `march`): only the ranges named by `offsets` are touched.
    totals (optional int32[>=2] device tensor): totals[1] receives the kept count.
      capacity=None: the single host sync of this call reads the whole `totals` tensor, so the caller gets totals[0] (e.g.
                     the marcher's count) for free; the python list is returned in place of `keep`.
      capacity=C   : NO host sync -- outputs have C entries (C >= the kept count, e.g. the marcher's capacity), the kept
                     count stays on the device (offsets'[R] and totals[1]).
    thre_cap (optional float32[1] device tensor): the threshold is min(alpha_thre, thre_cap[0]) -- mean(grid.occs) left on the device."""
    lib = _lib.load()
    thre_cap = _chk(thre_cap, torch.float32, "thre_cap", 1, allow_none=True)
    alphas = _chk(alphas, torch.float32, "alphas", 1)
    offsets = _chk(offsets, torch.int32, "offsets", 1)
    t_starts = _chk(t_starts, torch.float32, "t_starts", 1)
    t_ends = _chk(t_ends, torch.float32, "t_ends", 1)
    R, n, dev = offsets.numel() - 1, alphas.numel(), alphas.device
    keep = _alloc(pool, "vis_keep", n, torch.uint8, dev)
    kept = torch.empty((R,), dtype=torch.int32, device=dev)
    _lib.check(lib.angio_visibility_mask(_p(alphas), _p(offsets), R, float(early_stop_eps), float(alpha_thre), _p(keep), _p(kept),
                                         None, None, _p(thre_cap), _stream()), "angio_visibility_mask")
    host_totals = None
    if capacity is not None:
        new_offsets = exclusive_scan(kept, totals[1:2] if totals is not None else None)
        n2 = int(capacity)
    elif totals is not None:
        new_offsets = exclusive_scan(kept, totals[1:2])
        host_totals = totals.tolist()                      # the step's one host sync
        n2 = host_totals[1]
    else:
        new_offsets = exclusive_scan(kept)
        n2 = int(new_offsets[-1].item())
    ray_idx = _alloc(pool, "kept_idx", n2, torch.int32, dev)
    t0 = _alloc(pool, "kept_t0", n2, torch.float32, dev)
    t1 = _alloc(pool, "kept_t1", n2, torch.float32, dev)
    if n2 > 0:
        _lib.check(lib.angio_compact_samples(_p(keep), _p(offsets), _p(new_offsets), R, _p(t_starts), _p(t_ends), n2, _p(ray_idx), _p(t0),
                                             _p(t1), _stream()), "angio_compact_samples")
    return ray_idx, t0, t1, new_offsets, (keep if host_totals is None else host_totals)


def alphas_two_phase(desc, params, packed, precision, rays_o, rays_d, ray_idx, t_starts, t_ends, offsets, early_stop_eps, k0=32, timing=None,
                     pool=None):
    """alpha_fn over the marched samples with early ray termination (see angio_b200.h "Two-phase visibility pass"):
    phase A = the first k0 samples of every ray, phase B = the rest of the rays whose transmittance is still >= early_stop_eps
    after them.  Entries of the returned alphas behind a ray's termination point are undefined; visibility_compact never
    reads them.  No host sync: the list lengths stay on the device.  Returns (alphas, evaluated int32[2] device tensor)."""
    lib = _lib.load()
    if k0 % 32 != 0 or k0 <= 0:
        raise ValueError("k0 must be a positive multiple of 32 (the visibility kernel's chunk)")
    offsets = _chk(offsets, torch.int32, "offsets", 1)
    R, dev = offsets.numel() - 1, offsets.device
    cap = t_starts.numel()
    alphas = _alloc(pool, "vis_alphas", cap, torch.float32, dev)
    evaluated = torch.zeros((2,), dtype=torch.int32, device=dev)
    counts = torch.empty((R,), dtype=torch.int32, device=dev)
    alive = torch.empty((R,), dtype=torch.uint8, device=dev)
    kw = dict(rays_o=rays_o, rays_d=rays_d, ray_idx=ray_idx, t_starts=t_starts, t_ends=t_ends)
    for phase, (skip, limit, mask, size) in enumerate(((0, k0, None, min(cap, R * k0)), (k0, -1, alive, cap))):
        _lib.check(lib.angio_ray_segment_counts(_p(offsets), R, skip, limit, _p(mask), _p(counts), _stream()), "angio_ray_segment_counts")
        seg = exclusive_scan(counts, evaluated[phase:phase + 1])
        ids = _alloc(pool, "vis_ids%d" % phase, size, torch.int32, dev)
        if size > 0:
            _lib.check(lib.angio_ray_segment_ids(_p(offsets), _p(seg), R, skip, _p(ids), _stream()), "angio_ray_segment_ids")
            if timing is not None:                         # bench.py: CUDA events around each MLP launch + its device sample count
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            if precision == PREC_BF16:
                mlp_forward(desc, params, packed, OUT_ALPHA, precision, out=alphas, sample_idx=ids, n_dev=seg[R:R + 1], pool=pool, **kw)
            else:
                # fp32 check path (any MLP shape): the subset is gathered into exact-size arrays (one host read of its length)
                m = int(seg[R].item())
                if m > 0:
                    sel = ids[:m].long()
                    sub = mlp_forward(desc, params, packed, OUT_ALPHA, precision, rays_o=rays_o, rays_d=rays_d,
                                      ray_idx=ray_idx[sel].contiguous(), t_starts=t_starts[sel].contiguous(), t_ends=t_ends[sel].contiguous())
                    alphas[sel] = sub
            if timing is not None:
                ev1.record()
                timing.append((ev0, ev1, evaluated[phase:phase + 1]))
        if phase == 0:
            _lib.check(lib.angio_visibility_head(_p(alphas), _p(offsets), R, int(k0), float(early_stop_eps), _p(alive), _stream()),
                       "angio_visibility_head")
    return alphas, evaluated


def march_head(rays_o, rays_d, scene_aabb, roi_aabb, resolution, binary, near_plane, far_plane, step_size, k0=32, pool=None):
    """First k0 (<= 32) samples of every ray in ONE marching pass, packed without a scan (ray r: head_cnt[r] samples from slot
    head_base[r]; the order of the rays in memory is unspecified): returns (head_idx, head_t0, head_t1, head_cnt, head_base,
    head_total, t_resume, t_max).  Depends on the occupancy grid only, not on the model."""
    lib = _lib.load()
    if not 1 <= k0 <= 32:
        raise ValueError("k0 must be in 1..32")
    rays_o = _chk(rays_o, torch.float32, "ray_origins", 2)
    rays_d = _chk(rays_d, torch.float32, "ray_directions", 2)
    binary = _bin_u8(binary, resolution)
    aabb = _host6(scene_aabb, "scene_aabb")
    roi = _host6(roi_aabb, "roi_aabb")
    R, dev = rays_o.shape[0], rays_o.device
    h_idx = _alloc(pool, "lz_h_idx", R * k0, torch.int32, dev)
    h_t0 = _alloc(pool, "lz_h_t0", R * k0, torch.float32, dev)
    h_t1 = _alloc(pool, "lz_h_t1", R * k0, torch.float32, dev)
    h_cnt = torch.empty((R,), dtype=torch.int32, device=dev)
    h_base = torch.empty((R,), dtype=torch.int32, device=dev)
    h_total = torch.zeros((1,), dtype=torch.int32, device=dev)
    t_res = torch.empty((R,), dtype=torch.float32, device=dev)
    t_max = torch.empty((R,), dtype=torch.float32, device=dev)
    _lib.check(lib.angio_march_head(_p(rays_o), _p(rays_d), R, aabb.ctypes.data, roi.ctypes.data, int(resolution), _p(binary),
                                    float(near_plane), float(far_plane), float(step_size), int(k0), _p(h_idx), _p(h_t0), _p(h_t1), _p(h_cnt),
                                    _p(h_base), _p(h_total), _p(t_res), _p(t_max), _stream()), "angio_march_head")
    return h_idx, h_t0, h_t1, h_cnt, h_base, h_total, t_res, t_max


def march_filter_lazy(desc, params, packed, precision, rays_o, rays_d, scene_aabb, roi_aabb, resolution, binary, near_plane, far_plane,
                      step_size, early_stop_eps, alpha_thre, k0=32, totals=None, pool=None, timing=None, head=None, thre_cap=None):
    """acc_ray_marching (march -> alpha_fn -> visibility filter -> compaction) with LAZY marching and no host sync:

      head   the first k0 samples of every ray, one marching pass (no count / scan: warps reserve their slots atomically),
             alpha, visibility of the head, and which rays are still alive behind it;
      tail   only for those rays: count -> scan -> write from where the head stopped, alpha, visibility continuing the head's
             transmittance;
      then   one scan of the kept counts and one compaction of head + tail into the packed layout.

    The kept samples are bit-identical to marching every ray to the end and filtering afterwards (the reference's order); rays
    that are opaque after k0 samples -- or simply have no more -- never pay for the rest.  Returns (ray_idx, t_starts, t_ends,
    offsets) with capacity-sized arrays, the kept count in offsets[R]; totals (int32[>=2] device tensor) receives
    [samples marched (head + tail), samples kept].  `head` = the result of march_head for the same rays / grid / k0 when it was
    computed ahead of time (it does not depend on the model)."""
    lib = _lib.load()
    if precision != PREC_BF16:
        raise ValueError("march_filter_lazy: bf16 path only")
    thre_cap = _chk(thre_cap, torch.float32, "thre_cap", 1, allow_none=True)
    if not 1 <= k0 <= 32:
        raise ValueError("k0 must be in 1..32")
    rays_o = _chk(rays_o, torch.float32, "ray_origins", 2)
    rays_d = _chk(rays_d, torch.float32, "ray_directions", 2)
    binary = _bin_u8(binary, resolution)
    aabb = _host6(scene_aabb, "scene_aabb")
    roi = _host6(roi_aabb, "roi_aabb")
    R, dev = rays_o.shape[0], rays_o.device
    cap = march_capacity(R, near_plane, far_plane, step_size)
    i32, f32, u8 = torch.int32, torch.float32, torch.uint8
    # ---- head
    if head is None:
        head = march_head(rays_o, rays_d, aabb, roi, resolution, binary, near_plane, far_plane, step_size, k0=k0, pool=pool)
    h_idx, h_t0, h_t1, h_cnt, h_base, h_total, t_res, t_max = head
    if h_t0.numel() < R * k0 or h_cnt.numel() != R:
        raise ValueError("march_filter_lazy: `head` does not belong to these rays / this k0")
    h_alpha = _alloc(pool, "lz_h_alpha", R * k0, f32, dev)
    h_keep = _alloc(pool, "lz_h_keep", R * k0, u8, dev)
    h_kept = torch.empty((R,), dtype=i32, device=dev)
    t_end = torch.empty((R,), dtype=f32, device=dev)
    alive = torch.empty((R,), dtype=u8, device=dev)
    tail_total = torch.zeros((1,), dtype=i32, device=dev)
    if timing is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    s = Samples(R * k0, None, _p(rays_o), _p(rays_d), _p(h_idx), _p(h_t0), _p(h_t1), None, _p(h_total))
    _lib.check(lib.angio_mlp_forward(ctypes.byref(desc), _p(params), _p(packed), ctypes.byref(s), OUT_ALPHA, PREC_BF16, _p(h_alpha), None,
                                     None, 0, _stream()), "angio_mlp_forward")
    if timing is not None:
        ev1.record()
        timing.append((ev0, ev1, h_total))
    _lib.check(lib.angio_visibility_head_mask(_p(h_alpha), _p(h_cnt), _p(h_base), R, int(k0), float(early_stop_eps), float(alpha_thre),
                                              _p(h_keep), _p(h_kept), _p(t_end), _p(alive), _p(thre_cap), _stream()), "angio_visibility_head_mask")
    # ---- tail of the rays that are still alive (usually few; the kernels do nothing for the others)
    counts = torch.empty((R,), dtype=i32, device=dev)
    runs = _alloc(pool, "lz_runs", int(lib.angio_march_runs_bytes(R)), u8, dev)
    _lib.check(lib.angio_march_count(_p(rays_o), _p(rays_d), R, aabb.ctypes.data, roi.ctypes.data, int(resolution), _p(binary),
                                     float(near_plane), float(far_plane), float(step_size), _p(t_res), _p(t_max), _p(counts), _p(runs),
                                     _p(alive), _stream()), "angio_march_count")
    t_off = exclusive_scan(counts, tail_total)
    t_idx = _alloc(pool, "lz_t_idx", cap, i32, dev)
    t_t0 = _alloc(pool, "lz_t_t0", cap, f32, dev)
    t_t1 = _alloc(pool, "lz_t_t1", cap, f32, dev)
    t_alpha = _alloc(pool, "lz_t_alpha", cap, f32, dev)
    t_keep = _alloc(pool, "lz_t_keep", cap, u8, dev)
    _lib.check(lib.angio_march_write(_p(rays_o), _p(rays_d), R, roi.ctypes.data, int(resolution), _p(binary), float(step_size), _p(t_res),
                                     _p(t_max), _p(t_off), _p(runs), cap, _p(t_idx), _p(t_t0), _p(t_t1), _stream()), "angio_march_write")
    if timing is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    s = Samples(cap, None, _p(rays_o), _p(rays_d), _p(t_idx), _p(t_t0), _p(t_t1), None, _p(t_off[R:R + 1]))
    _lib.check(lib.angio_mlp_forward(ctypes.byref(desc), _p(params), _p(packed), ctypes.byref(s), OUT_ALPHA, PREC_BF16, _p(t_alpha), None,
                                     None, 0, _stream()), "angio_mlp_forward")
    if timing is not None:
        ev1.record()
        timing.append((ev0, ev1, tail_total))
    kept = torch.empty((R,), dtype=i32, device=dev)
    _lib.check(lib.angio_visibility_mask(_p(t_alpha), _p(t_off), R, float(early_stop_eps), float(alpha_thre), _p(t_keep), _p(kept), _p(t_end),
                                         _p(h_kept), _p(thre_cap), _stream()), "angio_visibility_mask")
    # ---- packed result
    new_off = exclusive_scan(kept, totals[1:2] if totals is not None else None)
    ray_idx = _alloc(pool, "kept_idx", cap, i32, dev)
    t0 = _alloc(pool, "kept_t0", cap, f32, dev)
    t1 = _alloc(pool, "kept_t1", cap, f32, dev)
    _lib.check(lib.angio_compact_head_tail(_p(h_keep), _p(h_cnt), _p(h_base), _p(h_t0), _p(h_t1), _p(t_keep), _p(t_off), _p(t_t0), _p(t_t1),
                                           _p(new_off), R, cap, _p(ray_idx), _p(t0), _p(t1), _stream()), "angio_compact_head_tail")
    if totals is not None:
        torch.add(h_total, tail_total, out=totals[0:1])
    return ray_idx, t0, t1, new_off


# ------------------------------------------------------------------------------------------------ composite
def composite_forward(logits, t_starts, t_ends, offsets, zero_mask=None):
    lib = _lib.load()
    logits = _chk(logits, torch.float32, "predictions", 1)
    t_starts = _chk(t_starts, torch.float32, "t_starts", 1)
    t_ends = _chk(t_ends, torch.float32, "t_ends", 1)
    offsets = _chk(offsets, torch.int32, "offsets", 1)
    zero_mask = _chk(zero_mask, torch.uint8, "zero_mask", 1, allow_none=True)
    R = offsets.numel() - 1
    pix = torch.empty((R,), dtype=torch.float32, device=logits.device)
    _lib.check(lib.angio_composite_forward(_p(logits), _p(t_starts), _p(t_ends), _p(offsets), R, _p(zero_mask), _p(pix), _stream()),
               "angio_composite_forward")
    return pix


def composite_backward(logits, t_starts, t_ends, offsets, pix, grad_pix, zero_mask=None):
    lib = _lib.load()
    grad_pix = _chk(grad_pix, torch.float32, "grad_pix", 1)
    R = offsets.numel() - 1
    g = torch.empty_like(logits)
    _lib.check(lib.angio_composite_backward(_p(logits), _p(t_starts), _p(t_ends), _p(offsets), R, _p(zero_mask), _p(pix), _p(grad_pix),
                                            _p(g), _stream()), "angio_composite_backward")
    return g


def composite_mse_fused(logits, t_starts, t_ends, offsets, target, n_rays_total=None, pool=None):
    """Returns (pix[R], grad_logits[n], loss_sum[1])."""
    lib = _lib.load()
    logits = _chk(logits, torch.float32, "predictions", 1)
    target = _chk(target, torch.float32, "target", 1)
    R = offsets.numel() - 1
    pix = torch.empty((R,), dtype=torch.float32, device=logits.device)
    g = _alloc(pool, "glogits", logits.numel(), torch.float32, logits.device)
    loss = torch.zeros((1,), dtype=torch.float32, device=logits.device)
    _lib.check(lib.angio_composite_mse_fused(_p(logits), _p(t_starts), _p(t_ends), _p(offsets), R, _p(target),
                                             int(n_rays_total or R), _p(pix), _p(g), _p(loss), _stream()), "angio_composite_mse_fused")
    return pix, g, loss


# ------------------------------------------------------------------------------------------------ MLP
def mlp_desc(enc, enc_basis, width, n_hidden) -> MlpDesc:
    return MlpDesc(int(enc), int(enc_basis), int(width), int(n_hidden))


def mlp_param_count(desc) -> int:
    n = int(_lib.load().angio_mlp_param_count(ctypes.byref(desc)))
    if n < 0:
        raise RuntimeError("angio_mlp_param_count: " + _lib.last_error())
    return n


def mlp_bf16_supported(desc) -> bool:
    return int(_lib.load().angio_mlp_packed_bytes(ctypes.byref(desc))) > 0


def mlp_pack(desc, params, packed=None):
    lib = _lib.load()
    params = _chk(params, torch.float32, "params", 1)
    nbytes = int(lib.angio_mlp_packed_bytes(ctypes.byref(desc)))
    if nbytes <= 0:
        raise RuntimeError("bf16 tensor-core path does not support this MLP shape: " + _lib.last_error())
    if packed is None or packed.numel() < nbytes:
        packed = torch.empty((nbytes,), dtype=torch.uint8, device=params.device)
    _lib.check(lib.angio_mlp_pack_weights(ctypes.byref(desc), _p(params), _p(packed), _stream()), "angio_mlp_pack_weights")
    return packed


def _samples(points=None, rays_o=None, rays_d=None, ray_idx=None, t_starts=None, t_ends=None, n_dev=None, sample_idx=None):
    if points is not None:
        points = _chk(points, torch.float32, "points", 2)
        if points.shape[1] != 3:
            raise ValueError("points must be [n, 3]")
        n = points.shape[0]
    else:
        rays_o = _chk(rays_o, torch.float32, "ray_origins", 2)
        rays_d = _chk(rays_d, torch.float32, "ray_directions", 2)
        ray_idx = _chk(ray_idx, torch.int32, "ray_idx", 1)
        n = ray_idx.numel()
    t_starts = _chk(t_starts, torch.float32, "t_starts", 1, allow_none=points is not None)
    t_ends = _chk(t_ends, torch.float32, "t_ends", 1, allow_none=points is not None)
    n_dev = _chk(n_dev, torch.int32, "n_dev", allow_none=True)
    sample_idx = _chk(sample_idx, torch.int32, "sample_idx", 1, allow_none=True)
    if sample_idx is not None:
        n = sample_idx.numel()
    s = Samples(n, _p(points), _p(rays_o), _p(rays_d), _p(ray_idx), _p(t_starts), _p(t_ends), _p(sample_idx), _p(n_dev))
    return s, n


_ITEMSIZE = {torch.float32: 4, torch.int32: 4, torch.uint8: 1, torch.int64: 8, torch.float64: 8}


class BufferPool:
    """Grow-only device scratch buffers (saved activations / workspaces / sample arrays) so steady-state steps never hit
    cudaMalloc.  A buffer that must grow doubles (the kept-sample count rises steadily while a model trains; a 1.3x policy
    re-allocated multi-GB buffers every few iterations) and the old block is released first."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, nbytes, device):
        b = self._bufs.get(key)
        if b is None or b.numel() < nbytes or b.device != device:
            self._bufs[key] = b = None                       # drop the old block before asking for the bigger one
            b = torch.empty((int(max(nbytes, 1)) * 2 + 256,), dtype=torch.uint8, device=device)
            self._bufs[key] = b
        return b

    def reserve(self, key, nbytes, device):
        """Exact-size allocation up front (worst-case preallocation of the sync-free training loop)."""
        b = self._bufs.get(key)
        if b is None or b.numel() < nbytes or b.device != device:
            self._bufs[key] = b = None
            self._bufs[key] = torch.empty((int(max(nbytes, 1)) + 256,), dtype=torch.uint8, device=device)
        return self._bufs[key]

    def typed(self, key, n, dtype, device):
        """n elements of dtype carved from the pooled buffer `key`."""
        item = _ITEMSIZE[dtype]
        return self.get(key, n * item, device)[:n * item].view(dtype)


def mlp_forward(desc, params, packed, out_mode, precision, saved=False, pool=None, out=None, **sample_kw):
    """Returns out[n] (and the saved-activation buffer when saved=True).  With sample_idx (an int32 index list into the
    ray-sample arrays) only those samples are evaluated and their results land at out[sample_idx[k]]; pass `out`
    (sized like t_starts) to collect several such calls in one array."""
    lib = _lib.load()
    params = _chk(params, torch.float32, "params", 1)
    s, n = _samples(**sample_kw)
    dev = params.device
    if out is None:
        n_out = sample_kw["t_starts"].numel() if sample_kw.get("sample_idx") is not None else n
        out = _alloc(pool if saved else None, "logits", n_out, torch.float32, dev)   # pooled only on the training path
    else:
        out = _chk(out, torch.float32, "out", 1)
    saved_buf = None
    if saved:
        sb = int(lib.angio_mlp_saved_bytes(ctypes.byref(desc), n, precision))
        if sb < 0:
            raise RuntimeError("angio_mlp_saved_bytes: " + _lib.last_error())
        saved_buf = pool.get("saved", sb, dev) if pool is not None else torch.empty((max(sb, 1),), dtype=torch.uint8, device=dev)
    wb = int(lib.angio_mlp_workspace_bytes(ctypes.byref(desc), n, precision, 0))
    if wb < 0:
        raise RuntimeError("angio_mlp_workspace_bytes: " + _lib.last_error())
    ws = pool.get("fwd_ws", wb, dev) if pool is not None else torch.empty((max(wb, 1),), dtype=torch.uint8, device=dev)
    _lib.check(lib.angio_mlp_forward(ctypes.byref(desc), _p(params), _p(packed), ctypes.byref(s), int(out_mode), int(precision),
                                     _p(out), _p(saved_buf), _p(ws), wb, _stream()), "angio_mlp_forward")
    return (out, saved_buf) if saved else out


def mlp_backward(desc, params, packed, saved_buf, grad_out, precision, grad_params=None, pool=None, **sample_kw):
    lib = _lib.load()
    params = _chk(params, torch.float32, "params", 1)
    grad_out = _chk(grad_out, torch.float32, "grad_out", 1)
    s, n = _samples(**sample_kw)
    if grad_out.numel() != n:
        raise ValueError("grad_out must have one element per sample")
    if grad_params is None:
        grad_params = torch.empty_like(params)
    wb = int(lib.angio_mlp_workspace_bytes(ctypes.byref(desc), n, precision, 1))
    if wb < 0:
        raise RuntimeError("angio_mlp_workspace_bytes: " + _lib.last_error())
    ws = pool.get("bwd_ws", wb, params.device) if pool is not None else torch.empty((max(wb, 1),), dtype=torch.uint8, device=params.device)
    _lib.check(lib.angio_mlp_backward(ctypes.byref(desc), _p(params), _p(packed), ctypes.byref(s), _p(saved_buf), _p(grad_out),
                                      int(precision), _p(grad_params), _p(ws), wb, _stream()), "angio_mlp_backward")
    return grad_params


# ------------------------------------------------------------------------------------------------ occupancy grid
def grid_cell_points(cells, jitter, roi_aabb, resolution):
    lib = _lib.load()
    jitter = _chk(jitter, torch.float32, "jitter", 2)
    cells = _chk(cells, torch.int64, "cells", 1, allow_none=True)
    n = jitter.shape[0]
    roi = _host6(roi_aabb, "roi_aabb")
    pts = torch.empty((n, 3), dtype=torch.float32, device=jitter.device)
    _lib.check(lib.angio_grid_cell_points(_p(cells), _p(jitter), n, roi.ctypes.data, int(resolution), _p(pts), _stream()),
               "angio_grid_cell_points")
    return pts


def grid_ema_update(occs, cells, occ, decay):
    lib = _lib.load()
    occs = _chk(occs, torch.float32, "occs", 1)
    occ = _chk(occ, torch.float32, "occ", 1)
    cells = _chk(cells, torch.int64, "cells", 1, allow_none=True)
    ws, wb = None, 0
    if cells is not None:
        wb = ((occs.numel() + 31) // 32) * 4
        ws = torch.empty((wb,), dtype=torch.uint8, device=occs.device)
    _lib.check(lib.angio_grid_ema_update(_p(occs), occs.numel(), _p(cells), _p(occ), occ.numel(), float(decay), _p(ws), wb, _stream()),
               "angio_grid_ema_update")


def grid_threshold(occs, occ_thre, binary_u8):
    """binary = occs > min(mean(occs), occ_thre).  Returns the device scalar mean(occs)."""
    lib = _lib.load()
    occs = _chk(occs, torch.float32, "occs", 1)
    binary_u8 = _chk(binary_u8, torch.uint8, "binary")
    mean = torch.empty((1,), dtype=torch.float32, device=occs.device)
    ws = torch.empty((4096,), dtype=torch.uint8, device=occs.device)
    _lib.check(lib.angio_grid_threshold(_p(occs), occs.numel(), float(occ_thre), _p(binary_u8), _p(mean), _p(ws), 4096, _stream()),
               "angio_grid_threshold")
    return mean


# ------------------------------------------------------------------------------------------------ optimiser
def adam_step(params, grads, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, active=None):
    """active: optional float32[1] device tensor; the step is skipped when it holds 0 (see angio_adam_step)."""
    lib = _lib.load()
    for t, nme in ((params, "params"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _chk(t, torch.float32, nme, 1)
    _chk(grads, torch.float32, "grads", 1)
    if grads.numel() < params.numel():
        raise ValueError("grads is smaller than params")
    active = _chk(active, torch.float32, "active", 1, allow_none=True)
    _lib.check(lib.angio_adam_step(_p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), float(lr), float(beta1),
                                   float(beta2), float(eps), int(step), float(grad_scale), _p(active), _stream()), "angio_adam_step")


def signal_peers(peer, tag):
    """Publish this rank's step tag to every peer (after its gradient is complete on the current stream)."""
    _lib.check(_lib.load().angio_signal_peers(peer.peer_flag_ptrs, peer.world, peer.rank, int(tag) & 0xFFFFFFFF, _stream()), "angio_signal_peers")


def adam_step_allreduce(params, peer, tag, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0,
                        active_index=-1, wait_stats=None):
    """Adam on the rank-ordered sum of all ranks' gradients, read straight from NVLink peer memory (see angio_b200.h)."""
    lib = _lib.load()
    for t, nme in ((params, "params"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _chk(t, torch.float32, nme, 1)
    _lib.check(lib.angio_adam_step_allreduce(_p(params), peer.peer_grad_ptrs[tag & 1], peer.world, _p(peer.flags), int(tag) & 0xFFFFFFFF,
                                             _p(exp_avg), _p(exp_avg_sq), params.numel(), float(lr), float(beta1), float(beta2), float(eps),
                                             int(step), float(grad_scale), int(active_index), _p(wait_stats), _stream()), "angio_adam_step_allreduce")


def project_volume(volume, bounds, rays_o, rays_d, depths, kind="ct"):
    """Ground-truth Beer-Lambert projection of an attenuation volume [X,Y,Z] along rays (phantomdata/helpers.py:192-224)."""
    lib = _lib.load()
    volume = _chk(volume, torch.float32, "volume", 3)
    rays_o = _chk(rays_o, torch.float32, "ray_origins", 2)
    rays_d = _chk(rays_d, torch.float32, "ray_directions", 2)
    depths = _chk(depths, torch.float32, "depths", 1)
    if kind not in ("ct", "sdf"):
        raise ValueError("kind must be 'ct' or 'sdf'")
    b = _host6(bounds, "bounds")
    n = rays_o.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=volume.device)
    _lib.check(lib.angio_project_volume(_p(volume), volume.shape[0], volume.shape[1], volume.shape[2], b.ctypes.data, _p(rays_o), _p(rays_d), n,
                                        _p(depths), depths.numel(), 1 if kind == "ct" else 0, _p(out), _stream()), "angio_project_volume")
    return out
