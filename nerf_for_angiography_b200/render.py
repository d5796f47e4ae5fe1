"""render_rays -- the reference's inline render sequence (/root/reference/nerf/run_nerf_acc.py:287-296, repeated at
:340-349 and visualization/visualization.py:335-345) as one call."""
import torch

from .nerf.nerf_helpers_acc import acc_ray_marching, acc_render_volume_density
from .nerf.nerf_helpers import get_predictions
from .model.CPPN import CPPN


def render_rays(model, grid, scene_aabb, ray_origins, ray_directions, depth_samples_per_ray, near_thresh, far_thresh,
                early_stop_eps=1e-2, alpha_thre=1e-3, chunksize=1024 * 128):
    """Returns (pix[n_rays], (ray_indices, t_starts, t_ends)).  Differentiable w.r.t. the model parameters."""
    n_rays = ray_origins.shape[0]
    with torch.no_grad():
        ray_indices, t_starts, t_ends = acc_ray_marching(model, grid, scene_aabb, ray_origins, ray_directions,
                                                         depth_samples_per_ray, near_thresh, far_thresh, early_stop_eps,
                                                         alpha_thre)
    if len(ray_indices) == 0:
        return torch.ones(n_rays, dtype=torch.float32, device=ray_origins.device), (ray_indices, t_starts, t_ends)
    if isinstance(model, CPPN):
        # midpoints are formed inside the MLP kernel (fuses run_nerf_acc.py:290-292 away)
        predictions = model.forward_samples(ray_origins.contiguous().float(), ray_directions.contiguous().float(),
                                            ray_indices._angio_idx32, t_starts.reshape(-1), t_ends.reshape(-1))
    else:
        idx = ray_indices.long()
        positions = ray_origins[idx] + ray_directions[idx] * (t_starts + t_ends) / 2.0
        predictions = get_predictions(model, positions, chunksize)
    pix, _ = acc_render_volume_density(predictions, ray_indices, t_starts, t_ends, n_rays, depth_samples_per_ray)
    return pix, (ray_indices, t_starts, t_ends)
