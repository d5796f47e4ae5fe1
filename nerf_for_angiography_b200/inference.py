"""Inference entry points of the reference's visualization script on the fused kernels
(/root/reference/visualization/visualization.py:209-229 volume query, :315-354 novel-view rendering), sharded across ranks:
views by contiguous ranges, the volume by slabs of its first axis; results gathered on rank 0."""
import numpy as np
import torch

from . import ops
from .distributed import gather_concat, shard_range, world
from .geometry import source_matrix
from .nerf.nerf_helpers_acc import acc_ray_marching, acc_render_volume_density


@torch.no_grad()
def render_projections(model, grid, scene_aabb, views, src_pt, img_width, img_height, focal_length, depth_samples_per_ray, near_thresh,
                       far_thresh, early_stop_eps=1e-2, alpha_thre=1e-3, binary_thresh=None, larm=0.0, translation=(0, 0, 0),
                       gather=True, shard=True):
    """Render novel views [(theta, phi), ...] through the hot path (visualization.py:315-354 for data_name == 'ct').
    Returns images [V, H, W] (and the 'binary' renders with sigma < binary_thresh zeroed when binary_thresh is given).
    With torch.distributed initialised every rank renders a contiguous slice of the views (shard=False: `views` is already
    this rank's share and all of it is rendered here)."""
    rank, ws = world() if shard else (0, 1)
    lo, hi = shard_range(len(views), rank, ws)
    dev = scene_aabb.device
    mats = np.stack([source_matrix(src_pt, th, ph, larm, translation) for th, ph in views[lo:hi]]) if hi > lo else np.zeros((0, 4, 4))
    cam = torch.from_numpy(mats).to(dev)
    n_rays = int(img_width) * int(img_height)
    imgs = torch.empty((hi - lo, int(img_height), int(img_width)), dtype=torch.float32, device=dev)
    bins = torch.empty_like(imgs) if binary_thresh is not None else None
    for v in range(hi - lo):
        o, d = ops.raygen(cam, int(img_width), int(img_height), float(focal_length), view=v)
        ri, ts, te = acc_ray_marching(model, grid, scene_aabb, o, d, depth_samples_per_ray, near_thresh, far_thresh, early_stop_eps,
                                      alpha_thre)
        if len(ri) == 0:
            imgs[v] = 1.0
            if bins is not None:
                bins[v] = 1.0
            continue
        pred = model.query(ops.OUT_LOGIT, rays_o=o, rays_d=d, ray_idx=ri._angio_idx32, t_starts=ts.reshape(-1), t_ends=te.reshape(-1))
        pix, _ = acc_render_volume_density(pred, ri, ts, te, n_rays, depth_samples_per_ray)
        imgs[v] = pix.view(int(img_height), int(img_width))
        if bins is not None:
            zero_idx = torch.where(torch.sigmoid(pred) < binary_thresh)
            pb, _ = acc_render_volume_density(pred, ri, ts, te, n_rays, depth_samples_per_ray, zero_idx)
            bins[v] = pb.view(int(img_height), int(img_width))
    if gather and ws > 1:
        imgs = gather_concat(imgs)
        bins = gather_concat(bins) if bins is not None else None
    return (imgs, bins) if binary_thresh is not None else imgs


@torch.no_grad()
def query_volume(model, t, grid=None, chunk=1 << 24, gather=True):
    """sigma = sigmoid(model(x)) on the lattice np.meshgrid(t, t, t) (default 'xy' indexing, as the reference builds it at
    visualization.py:209): out[i, j, k] is the value at (t[j], t[i], t[k]).  If `grid` is given, cells the occupancy
    grid marks empty are returned as 0 without evaluating the model.  Ranks split the first axis into slabs."""
    rank, ws = world()
    dev = model._flat.device if model._flat is not None else torch.device("cuda")
    t = torch.as_tensor(t, dtype=torch.float32, device=dev)
    n = t.numel()
    lo, hi = shard_range(n, rank, ws)
    out = torch.empty((hi - lo, n, n), dtype=torch.float32, device=dev)
    rows_per_chunk = max(1, int(chunk) // (n * n))
    for i0 in range(lo, hi, rows_per_chunk):
        i1 = min(hi, i0 + rows_per_chunk)
        yy, xx, zz = torch.meshgrid(t[i0:i1], t, t, indexing="ij")        # first axis carries y ('xy' meshgrid)
        pts = torch.stack([xx, yy, zz], dim=-1).reshape(-1, 3).contiguous()
        sig = model.query(ops.OUT_SIGMA, points=pts)
        if grid is not None:
            sig = sig * grid.query_occ(pts)
        out[i0 - lo:i1 - lo] = sig.view(i1 - i0, n, n)
    if gather and ws > 1:
        out = gather_concat(out)
    return out
