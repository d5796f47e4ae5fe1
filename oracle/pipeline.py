"""oracle/pipeline.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference's hot-path glue restated on torch-CPU fp32, function for function:

  acc_ray_marching           /root/reference/nerf/nerf_helpers_acc.py:10-31
  acc_render_volume_density  /root/reference/nerf/nerf_helpers_acc.py:45-63
  acc_update_n_step          /root/reference/nerf/nerf_helpers_acc.py:65-78
  get_predictions            /root/reference/nerf/nerf_helpers.py:24-45
  midpoint positions         /root/reference/nerf/run_nerf_acc.py:290-292
  train_step                 /root/reference/nerf/run_nerf_acc.py:284-307,323-328 (Adam via torch.optim)

``model`` everywhere is a callable x[S,3] -> logit[S,1] (e.g. functools.partial(cppn_forward, params)).
"""
import numpy as np
import torch

from . import nerfacc_ref


def midpoints(rays_o, rays_d, ray_indices, t_starts, t_ends):
    idx = torch.as_tensor(ray_indices).long()
    return rays_o[idx] + rays_d[idx] * (t_starts + t_ends) / 2.0      # run_nerf_acc.py:290-292


def get_predictions(model, pts, chunksize):
    outs = [model(pts[i:i + chunksize]) for i in range(0, pts.shape[0], chunksize)]
    return torch.cat(outs, dim=0) if outs else torch.zeros((0, 1), dtype=pts.dtype)


def acc_ray_marching(model, grid, scene_aabb, rays_o, rays_d, depth_samples_per_ray, near, far,
                     early_stop_eps=1e-2, alpha_thre=1e-3, return_prefilter=False):
    rays_o = torch.as_tensor(rays_o, dtype=torch.float32)
    rays_d = torch.as_tensor(rays_d, dtype=torch.float32)

    def alpha_fn(t_starts, t_ends, ray_indices):            # nerf_helpers_acc.py:11-25
        with torch.no_grad():
            ts, te = torch.from_numpy(t_starts), torch.from_numpy(t_ends)
            pos = midpoints(rays_o, rays_d, ray_indices, ts, te)
            sig = torch.sigmoid(model(pos))
            alphas = torch.exp(-sig * (te - ts))
            return (1 - alphas).numpy()

    step = (far - near) / depth_samples_per_ray              # nerf_helpers_acc.py:27
    return nerfacc_ref.ray_marching(rays_o.numpy(), rays_d.numpy(), np.asarray(scene_aabb, np.float32), grid,
                                    alpha_fn, near, far, early_stop_eps, alpha_thre, step,
                                    return_prefilter=return_prefilter)


def acc_render_volume_density(predictions, ray_indices, t_starts, t_ends, n_rays, zero_mask=None):
    """Differentiable torch restatement; scatter_mul == out.scatter_reduce_(prod)."""
    dists = t_ends - t_starts
    sig = torch.sigmoid(predictions)
    if zero_mask is not None:
        sig = torch.where(zero_mask.reshape(sig.shape), torch.zeros_like(sig), sig)
    alphas = torch.exp(-sig * dists)
    index = torch.as_tensor(ray_indices).long()[:, None]
    out = torch.ones((n_rays, 1), dtype=alphas.dtype)
    out = out.scatter_reduce(0, index, alphas, reduce="prod", include_self=True)
    return out.squeeze(-1).float()


def acc_update_n_step(grid, model, step, occ_thre=1e-2, **kw):
    def occ_eval_fn(x):
        with torch.no_grad():
            return torch.sigmoid(model(torch.from_numpy(np.asarray(x, np.float32)))).numpy()
    grid.every_n_step(step, occ_eval_fn, occ_thre=occ_thre, **kw)
    return grid


def render_rays(model, grid, scene_aabb, rays_o, rays_d, n_steps, near, far, early_stop_eps, alpha_thre,
                chunksize=131072):
    """run_nerf_acc.py:287-296 in one call.  Returns (pix[R], (ray_indices, t_starts, t_ends))."""
    rays_o = torch.as_tensor(rays_o, dtype=torch.float32)
    rays_d = torch.as_tensor(rays_d, dtype=torch.float32)
    ri, ts, te = acc_ray_marching(model, grid, scene_aabb, rays_o, rays_d, n_steps, near, far,
                                  early_stop_eps, alpha_thre)
    ts_t, te_t = torch.from_numpy(ts), torch.from_numpy(te)
    n_rays = rays_o.shape[0]
    if len(ri) == 0:
        return torch.ones(n_rays), (ri, ts, te)
    pos = midpoints(rays_o, rays_d, ri, ts_t, te_t)
    pred = get_predictions(model, pos, chunksize)
    pix = acc_render_volume_density(pred, ri, ts_t, te_t, n_rays)
    return pix, (ri, ts, te)


def lr_at(n_iter, base_lr=1e-4, decay_rate=0.1, decay_steps=500000):
    return base_lr * (decay_rate ** (n_iter / decay_steps))     # run_nerf_acc.py:323-328
