"""Drop-in for the two hot-path functions of /root/reference/nerf/nerf_helpers.py."""
import torch


def get_minibatches(inputs, chunksize=1024 * 8):
    return [inputs[i:i + chunksize] for i in range(0, inputs.shape[0], chunksize)]


def get_predictions(model, flattened_query_points, chunksize, target_img_idx=None):
    """/root/reference/nerf/nerf_helpers.py:31-45.  The fused kernel streams sample tiles itself, so the chunk
    loop is kept only for generic models; a CPPN is evaluated in one launch (same values)."""
    if target_img_idx:
        raise NotImplementedError("target_img_idx (per-image index channel) is never enabled by the reference driver")
    from ..model.CPPN import CPPN
    if isinstance(model, CPPN):
        return model(flattened_query_points)
    preds = [model(b) for b in get_minibatches(flattened_query_points, chunksize=chunksize)]
    return torch.cat(preds, dim=0)


def sample_pixel_rays(ray_pool, img_sample_size, device=None, weights=None, unseen=False, generator=None):
    """/root/reference/nerf/nerf_helpers.py:137-150 on a device-resident RayPool (data.RayPool) instead of a pandas
    DataFrame: weighted sampling WITHOUT replacement over all rays of all views, then rays generated on the fly.
    Returns [origins[R,3], directions[R,3], pixel_values[R]]."""
    o, d, pix = ray_pool.sample(img_sample_size, weights=weights, generator=generator)
    return [o, d, None if unseen else pix]
