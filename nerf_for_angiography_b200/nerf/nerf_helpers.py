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


def sample_pixel_rays(train_ray_df, img_sample_size, device=None, weights=None, unseen=False, generator=None):
    """/root/reference/nerf/nerf_helpers.py:137-150: weighted sampling WITHOUT replacement of `img_sample_size` rays over all
    rays of all views, shuffled.  Returns [origins[R,3], directions[R,3], pixel_values[R] (None if unseen)].

    `train_ray_df` may be
      * the reference's pandas ray DataFrame (columns ray_origins_{x,y,z} / ray_directions_{x,y,z} or the list columns
        'ray_origins' / 'ray_directions', 'pixel_value', weight columns) with `weights` = a column name or None, exactly as the
        driver calls it (run_nerf_acc.py:277).  It is copied to the device once (cached on the frame) and sampled there;
      * a device-resident pool (`data.RayPool`: rays regenerated on the fly from (view, x, y), or `data.ExplicitRayPool`)."""
    from ..data import ExplicitRayPool, RayPool
    pool = train_ray_df
    if not isinstance(pool, (RayPool, ExplicitRayPool)):
        cache = train_ray_df.attrs.setdefault("_angio_pool", {})
        key = str(device or "cuda")
        if key not in cache:
            wcols = tuple(c for c in train_ray_df.columns if c not in ("image_id", "pixel_value", "x_position", "y_position", "ray_origins",
                                                                     "ray_directions") and not c.startswith(("ray_origins_", "ray_directions_")))
            cache[key] = ExplicitRayPool.from_dataframe(train_ray_df, device=device or "cuda", weight_columns=wcols)
        pool = cache[key]
    if isinstance(pool, RayPool):
        # a RayPool carries ONE weight image (the reference's distance_pixel_value): a column name selects it, None = uniform
        weights = None if isinstance(weights, str) else ("uniform" if weights is None else weights)
    o, d, pix = pool.sample(img_sample_size, weights=weights, generator=generator)
    return [o, d, None if unseen else pix]
