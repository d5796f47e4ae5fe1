"""Drop-in for /root/reference/nerf/nerf_helpers_acc.py: same function names, arguments and return values,
running on the sm_100a kernels (no nerfacc CUDA, no torch_scatter)."""
import torch

from .. import ops
from ..nerfacc import ray_marching, _is_fused_model


def acc_ray_marching(radiance_field, grid, scene_aabb, ray_origins, ray_directions, depth_samples_per_ray, near_thresh,
                     far_thresh, early_stop_eps=1e-2, alpha_thre=1e-3):
    """/root/reference/nerf/nerf_helpers_acc.py:10-31."""
    def alpha_fn(t_starts, t_ends, ray_indices):
        # generic models: the reference closure (positions materialised with torch ops, model called as is)
        t_origins = ray_origins[ray_indices]
        t_dirs = ray_directions[ray_indices]
        positions = t_origins + t_dirs * (t_starts + t_ends) / 2.0
        sigmas = torch.sigmoid(radiance_field(positions))
        return 1 - torch.exp(-sigmas * (t_ends - t_starts))

    render_step_size = (far_thresh - near_thresh) / depth_samples_per_ray
    fused = radiance_field if _is_fused_model(radiance_field) else None
    return ray_marching(ray_origins, ray_directions, scene_aabb=scene_aabb, grid=grid, alpha_fn=alpha_fn,
                        near_plane=near_thresh, far_plane=far_thresh, early_stop_eps=early_stop_eps, alpha_thre=alpha_thre,
                        render_step_size=render_step_size, radiance_field=fused)


def packed_offsets(ray_indices, n_rays):
    """Segment offsets [n_rays+1] int32 of a ray-sorted index vector (attached by ray_marching when available)."""
    off = getattr(ray_indices, "_angio_offsets", None)
    if off is not None and off.numel() == n_rays + 1:
        return off
    bounds = torch.arange(n_rays + 1, device=ray_indices.device, dtype=ray_indices.dtype)
    return torch.searchsorted(ray_indices.contiguous(), bounds).to(torch.int32)


class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predictions, t_starts, t_ends, offsets, zero_mask):
        pix = ops.composite_forward(predictions, t_starts, t_ends, offsets, zero_mask)
        ctx.save_for_backward(predictions, t_starts, t_ends, offsets, pix)
        ctx.zero_mask = zero_mask
        return pix

    @staticmethod
    def backward(ctx, grad_pix):
        predictions, t_starts, t_ends, offsets, pix = ctx.saved_tensors
        g = ops.composite_backward(predictions, t_starts, t_ends, offsets, pix, grad_pix.contiguous().float(), ctx.zero_mask)
        return g, None, None, None, None


def acc_render_volume_density(predictions, ray_indices, t_starts, t_ends, n_rays, depth_samples_per_ray, zero_idx=[]):
    """/root/reference/nerf/nerf_helpers_acc.py:45-63.  Returns (pix[n_rays] float32, None)."""
    n = predictions.shape[0]
    if predictions.dim() == 2 and predictions.shape[-1] != 1:
        raise NotImplementedError("acc_render_volume_density: single-channel attenuation only")
    if n > 1 and getattr(ray_indices, "_angio_offsets", None) is None:
        if not bool((ray_indices[1:] >= ray_indices[:-1]).all()):
            raise NotImplementedError("acc_render_volume_density: ray_indices must be sorted by ray (packed layout)")
    offsets = packed_offsets(ray_indices, int(n_rays))
    zero_mask = None
    if len(zero_idx) > 0:
        zero_mask = torch.zeros(n, dtype=torch.uint8, device=predictions.device)
        idx = zero_idx[0] if isinstance(zero_idx, (tuple, list)) else zero_idx
        if isinstance(idx, torch.Tensor) and idx.dtype == torch.bool:
            zero_mask[idx.reshape(-1)] = 1
        else:
            zero_mask[idx] = 1
    pix = _Composite.apply(predictions.reshape(-1).contiguous().float(), t_starts.reshape(-1).contiguous().float(),
                           t_ends.reshape(-1).contiguous().float(), offsets, zero_mask)
    return pix, None


def acc_update_n_step(acc_grid, radiance_field, step, occ_thre=1e-2, inverse=False):
    """/root/reference/nerf/nerf_helpers_acc.py:65-78 (occupancy = sigmoid(MLP(x)), not sigma*dt)."""
    if _is_fused_model(radiance_field):
        def occ_eval_fn(x):
            return radiance_field.query(ops.OUT_SIGMA, points=x)
    else:
        def occ_eval_fn(x):
            return torch.sigmoid(radiance_field(x))
    acc_grid.every_n_step(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre)
    return acc_grid
