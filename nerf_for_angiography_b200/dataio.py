"""Reference data contract: the two ';'-separated CSV files the phantom scripts write
(/root/reference/phantomdata/cttoray.py:271-308, sdftoray.py:177-208) and the training driver consumes
(/root/reference/nerf/run_nerf_acc.py:82-124) -- plus the step from those DataFrames to the device-resident RayPool.

proj CSV  df-{file_name}-{binary_str}-cttoproj.csv : one row per projection; columns
    image_id, theta, phi, larm, [theta_shift, phi_shift, larm_shift,] translation_{x,y,z}, tform_cam2world (4x4 list as
    text), [unshifted_tform_cam2world,] image_data, image_distance_data, org_img_width, org_img_height, focal_length,
    near_thresh, far_thresh, depth_sample, [grid_scaling_factor,] depth_values, src_pt_z
ray CSV   df-rays-{file_name}-{binary_str}-{img_height}.csv : one row per ray; columns
    image_id, pixel_value, distance_pixel_value, x_position, y_position, ray_origins_{x,y,z}, ray_directions_{x,y,z}
image_id = f'{theta}-{phi}' with '.' -> ',' (cttoray.py:191); the LAST projection row is the test view (run_nerf_acc.py:85).

`load_data` is called by the reference driver (run_nerf_acc.py:82) but is missing upstream; the version here returns the four
values the driver unpacks.  Host-side pandas code: this is the data format either side of the hot path, not the hot path.
"""
import ast
import os

import numpy as np


def image_id_of(theta, phi) -> str:
    """cttoray.py:191"""
    return f"{theta}-{phi}".replace(".", ",")


def proj_csv_name(file_name, binary_str):
    return f"df-{file_name}-{binary_str}-cttoproj.csv"


def ray_csv_name(file_name, binary_str, img_height):
    return f"df-rays-{file_name}-{binary_str}-{img_height}.csv"


def write_reference_csvs(folder, file_name, views, cam2world, images, dist_images, rays_o, rays_d, focal_length, near_thresh,
                         far_thresh, depth_samples, src_pt_z, binary_str="nonbinary", larm=0.0, translations=None):
    """Write one dataset in the reference's CSV layout.

    views: [(theta, phi)] with the test view LAST; cam2world [V,4,4] float64; images / dist_images [V,H,W];
    rays_o / rays_d [V,H,W,3] float64 as produced by get_ray_values (pixel order [row=y, col=x]).  Returns the two paths."""
    import pandas as pd
    os.makedirs(folder, exist_ok=True)
    V, H, W = np.asarray(images).shape
    images = np.asarray(images, dtype=np.float64)
    dist_images = np.asarray(dist_images, dtype=np.float64)
    translations = np.zeros((V, 3)) if translations is None else np.asarray(translations, dtype=np.float64)
    depth_values = np.linspace(near_thresh, far_thresh, int(depth_samples)).tolist()
    proj = pd.DataFrame({
        "image_id": [image_id_of(t, p) for t, p in views],
        "theta": [float(t) for t, _ in views], "phi": [float(p) for _, p in views], "larm": [float(larm)] * V,
        "translation_x": translations[:, 0], "translation_y": translations[:, 1], "translation_z": translations[:, 2],
        "tform_cam2world": [np.asarray(m, dtype=np.float64).tolist() for m in cam2world],
        "image_data": [im.tolist() for im in images], "image_distance_data": [im.tolist() for im in dist_images],
        "org_img_width": [W] * V, "org_img_height": [H] * V, "focal_length": [float(focal_length)] * V,
        "near_thresh": [float(near_thresh)] * V, "far_thresh": [float(far_thresh)] * V, "depth_sample": [int(depth_samples)] * V,
        "depth_values": [depth_values] * V, "src_pt_z": [float(src_pt_z)] * V})
    proj_path = os.path.join(folder, proj_csv_name(file_name, binary_str))
    proj.to_csv(proj_path, sep=";")
    jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")      # row (y), column (x): helpers.py:162-166
    rays_o = np.asarray(rays_o, dtype=np.float64).reshape(V, H * W, 3)
    rays_d = np.asarray(rays_d, dtype=np.float64).reshape(V, H * W, 3)
    ray = pd.DataFrame({
        "image_id": np.repeat([image_id_of(t, p) for t, p in views], H * W),
        "pixel_value": images.reshape(-1), "distance_pixel_value": dist_images.reshape(-1),
        "x_position": np.tile(ii.reshape(-1), V), "y_position": np.tile(jj.reshape(-1), V),
        "ray_origins_x": rays_o[..., 0].reshape(-1), "ray_origins_y": rays_o[..., 1].reshape(-1), "ray_origins_z": rays_o[..., 2].reshape(-1),
        "ray_directions_x": rays_d[..., 0].reshape(-1), "ray_directions_y": rays_d[..., 1].reshape(-1),
        "ray_directions_z": rays_d[..., 2].reshape(-1)})
    ray_path = os.path.join(folder, ray_csv_name(file_name, binary_str, H))
    ray.to_csv(ray_path, sep=";")
    return proj_path, ray_path


def read_reference_csvs(proj_path, ray_path):
    """(proj_df, ray_df) indexed the way the driver uses them: proj_df.index = image_id (run_nerf_acc.py:85-86)."""
    import pandas as pd
    proj = pd.read_csv(proj_path, sep=";", index_col=0, float_precision="round_trip")   # the reference's rays are float64 text
    ray = pd.read_csv(ray_path, sep=";", index_col=0, float_precision="round_trip")
    for col in ("tform_cam2world", "unshifted_tform_cam2world", "image_data", "image_distance_data", "depth_values"):
        if col in proj.columns:
            proj[col] = proj[col].map(ast.literal_eval)              # lists were serialised as text
    proj = proj.set_index("image_id", drop=False)
    return proj, ray


def load_data(data_name, file_name, unseen=False, binary=False, data_size=None, step_size=None, data_root="data"):
    """The function run_nerf_acc.py:82 calls: -> (proj_df, ray_df, store_folder_name, unseen_ray_df).

    Files are looked up under {data_root}/{data_name}/.  `unseen`: the rays of views listed in an optional
    df-rays-...-unseen CSV next to the training rays (else an empty frame).  data_size / step_size only name the output folder,
    as in the reference's folder convention."""
    import glob

    import pandas as pd
    folder = os.path.join(data_root, data_name)
    binary_str = "binary" if binary else "nonbinary"
    proj_path = os.path.join(folder, proj_csv_name(file_name, binary_str))
    cands = sorted(p for p in glob.glob(os.path.join(folder, f"df-rays-{file_name}-{binary_str}-*.csv")) if not p.endswith("-unseen.csv"))
    if not os.path.exists(proj_path) or not cands:
        raise FileNotFoundError(f"no reference dataset '{file_name}' ({binary_str}) under {folder}")
    proj_df, ray_df = read_reference_csvs(proj_path, cands[0])
    unseen_ray_df = pd.DataFrame(columns=ray_df.columns)
    if unseen:
        up = cands[0][:-4] + "-unseen.csv"
        if os.path.exists(up):
            unseen_ray_df = pd.read_csv(up, sep=";", index_col=0)
    store_folder_name = os.path.join("cases", data_name, f"{file_name}-{binary_str}" + (f"-{data_size}" if data_size is not None else "")
                                     + (f"-{step_size}" if step_size is not None else ""))
    return proj_df, ray_df, store_folder_name, unseen_ray_df


def pool_from_dataframes(proj_df, ray_df, device="cuda", validate=256):
    """Device-resident RayPool (cam2world, pixels, sampling weights) from the reference DataFrames.  A ray is regenerated on
    the fly from (view, x, y), so the CSV's precomputed origins / directions are only used to VALIDATE that the matrices
    reproduce them (`validate` random rays, 1e-6 relative) -- e.g. a CSV written with shifted matrices would fail here."""
    import torch

    from . import ops
    from .data import RayPool
    ids = list(proj_df.index)
    V = len(ids)
    W, H = int(proj_df["org_img_width"].iloc[0]), int(proj_df["org_img_height"].iloc[0])
    focal = float(proj_df["focal_length"].iloc[0])
    cam = np.stack([np.asarray(m, dtype=np.float64).reshape(4, 4) for m in proj_df["tform_cam2world"]])
    view_of = {v: k for k, v in enumerate(ids)}
    v_idx = ray_df["image_id"].map(view_of).to_numpy()
    if np.isnan(v_idx.astype(np.float64)).any():
        raise ValueError("ray CSV names an image_id the projection CSV does not contain")
    v_idx = v_idx.astype(np.int64)
    x = ray_df["x_position"].to_numpy().astype(np.int64)
    y = ray_df["y_position"].to_numpy().astype(np.int64)
    if len(ray_df) != V * H * W:
        raise ValueError(f"ray CSV holds {len(ray_df)} rays, expected {V}x{H}x{W}")
    pix = np.zeros((V, H, W), np.float32)
    wts = np.zeros((V, H, W), np.float32)
    pix[v_idx, y, x] = ray_df["pixel_value"].to_numpy().astype(np.float32)
    wts[v_idx, y, x] = ray_df["distance_pixel_value"].to_numpy().astype(np.float32)
    cam_t = torch.from_numpy(cam).to(device)
    pool = RayPool(cam_t, torch.from_numpy(pix).to(device), focal, torch.from_numpy(wts).to(device))
    if validate:
        rng = np.random.default_rng(0)
        sel = rng.choice(len(ray_df), size=min(int(validate), len(ray_df)), replace=False)
        flat = torch.from_numpy((v_idx[sel] * H + y[sel]) * W + x[sel]).to(device)
        o, d = ops.raygen_flat(cam_t, flat, W, H, focal)
        o_csv = ray_df[["ray_origins_x", "ray_origins_y", "ray_origins_z"]].to_numpy()[sel].astype(np.float32)
        d_csv = ray_df[["ray_directions_x", "ray_directions_y", "ray_directions_z"]].to_numpy()[sel].astype(np.float32)
        if not (np.allclose(o.cpu().numpy(), o_csv, rtol=1e-6, atol=1e-4) and np.allclose(d.cpu().numpy(), d_csv, rtol=1e-6, atol=1e-6)):
            raise ValueError("the CSV's precomputed rays are not the cone-beam rays of its tform_cam2world matrices")
    info = dict(focal=focal, near=float(proj_df["near_thresh"].iloc[0]), far=float(proj_df["far_thresh"].iloc[0]),
                depth_samples=int(proj_df["depth_sample"].iloc[0]), src_pt_z=float(proj_df["src_pt_z"].iloc[0]), views=ids)
    return pool, info
