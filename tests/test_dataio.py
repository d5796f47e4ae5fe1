"""The reference's CSV data contract (cttoray.py:271-308 -> run_nerf_acc.py:82-124): write / read round trip on the CPU,
and (GPU) the RayPool built from the DataFrames reproduces the CSV's precomputed rays."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry as ogeo


def _tiny_dataset(tmp_path, W=6, H=5):
    from nerf_for_angiography_b200 import dataio
    views = [(0.0, 0.0), (45.0, 0.0), (135.0, 135.0)]                 # test view last
    src = [0.0, 0.0, 1500.0]
    focal = 7.5 * W
    mats, os_, ds_ = [], [], []
    for th, ph in views:
        o, d, M = ogeo.get_ray_values(th, ph, 0.0, src, W, H, focal)
        mats.append(M); os_.append(o); ds_.append(d)
    rng = np.random.default_rng(0)
    images = rng.random((len(views), H, W))
    dist = rng.random((len(views), H, W)) + 0.1
    folder = os.path.join(tmp_path, "data", "phantom")
    paths = dataio.write_reference_csvs(folder, "case-a", views, np.stack(mats), images, dist, np.stack(os_), np.stack(ds_), focal, 1400.0,
                                        1600.0, 300, 1500.0)
    return dataio, views, np.stack(mats), images, dist, np.stack(os_), np.stack(ds_), paths, W, H, focal


def test_csv_contract_round_trip(tmp_path):
    dataio, views, mats, images, dist, o, d, paths, W, H, focal = _tiny_dataset(str(tmp_path))
    assert os.path.basename(paths[0]) == "df-case-a-nonbinary-cttoproj.csv" and os.path.basename(paths[1]) == f"df-rays-case-a-nonbinary-{H}.csv"
    head = open(paths[1]).readline().strip().split(";")
    assert head[1:] == ["image_id", "pixel_value", "distance_pixel_value", "x_position", "y_position", "ray_origins_x", "ray_origins_y",
                        "ray_origins_z", "ray_directions_x", "ray_directions_y", "ray_directions_z"]
    proj_df, ray_df, store, unseen = dataio.load_data("phantom", "case-a", False, False, 100, 1, data_root=os.path.join(str(tmp_path), "data"))
    assert list(proj_df.index) == ["0,0-0,0", "45,0-0,0", "135,0-135,0"]                       # '.' -> ',' (cttoray.py:191)
    assert len(unseen) == 0 and "case-a" in store
    # what run_nerf_acc.py:85-124 reads
    test_id = proj_df.index[-1]
    test_rays = ray_df[ray_df["image_id"] == test_id]
    assert len(test_rays) == W * H
    assert int(test_rays["x_position"].max()) + 1 == W and int(test_rays["y_position"].max()) + 1 == H
    assert float(proj_df["focal_length"].iloc[0]) == focal and int(proj_df["depth_sample"].iloc[0]) == 300
    assert np.array_equal(np.asarray(proj_df["tform_cam2world"].iloc[1]), mats[1])              # float64 repr round-trips exactly
    yy, xx = test_rays["y_position"].to_numpy(), test_rays["x_position"].to_numpy()
    assert np.array_equal(test_rays["pixel_value"].to_numpy(), images[2][yy, xx])
    assert np.array_equal(test_rays["ray_directions_x"].to_numpy(), d[2][yy, xx, 0])
    with pytest.raises(FileNotFoundError):
        dataio.load_data("phantom", "missing", data_root=os.path.join(str(tmp_path), "data"))


@pytest.mark.gpu
def test_pool_from_reference_dataframes(tmp_path):
    dataio, views, mats, images, dist, o, d, paths, W, H, focal = _tiny_dataset(str(tmp_path))
    proj_df, ray_df = dataio.read_reference_csvs(*paths)
    pool, info = dataio.pool_from_dataframes(proj_df, ray_df, device="cuda")
    assert (pool.n_views, pool.img_h, pool.img_w) == (3, H, W) and info["near"] == 1400.0 and info["far"] == 1600.0
    assert np.array_equal(pool.pixels.cpu().numpy(), images.astype(np.float32))
    assert np.array_equal(pool.weights.cpu().numpy(), dist.astype(np.float32))
    ro, rd, pix = pool.rays_of_view(2)
    assert np.array_equal(rd.cpu().numpy(), d[2].reshape(-1, 3).astype(np.float32))             # the test view's rays, bit for bit
    # a CSV whose matrices do not generate its rays is rejected
    bad = proj_df.copy()
    bad["tform_cam2world"] = [np.eye(4).tolist()] * 3
    with pytest.raises(ValueError):
        dataio.pool_from_dataframes(bad, ray_df, device="cuda")


@pytest.mark.gpu
def test_sample_pixel_rays_accepts_the_reference_dataframe(tmp_path):
    """run_nerf_acc.py:277 calls sample_pixel_rays(train_ray_df, n, device, weights=<column name>): same call, DataFrame in,
    device tensors out, rows drawn without replacement and consistent with the frame."""
    import nerf_for_angiography_b200 as A
    dataio, views, mats, images, dist, o, d, paths, W, H, focal = _tiny_dataset(str(tmp_path))
    proj_df, ray_df = dataio.read_reference_csvs(*paths)
    g = torch.Generator(device="cuda").manual_seed(0)
    bo, bd, bp = A.sample_pixel_rays(ray_df, 40, "cuda", weights="distance_pixel_value", generator=g)
    assert bo.shape == (40, 3) and bd.shape == (40, 3) and bp.shape == (40,) and bo.is_cuda
    rows = ray_df[["ray_directions_x", "ray_directions_y", "ray_directions_z"]].to_numpy().astype(np.float32)
    pix = ray_df["pixel_value"].to_numpy().astype(np.float32)
    hit = [int(np.flatnonzero((rows == r).all(1) & (pix == p))[0]) for r, p in zip(bd.cpu().numpy(), bp.cpu().numpy())]
    assert len(set(hit)) == 40                                                                # 40 distinct rows of the frame
    assert "_angio_pool" in ray_df.attrs                                                      # copied to the device once
    bo2, bd2, bp2 = A.sample_pixel_rays(ray_df, 90, "cuda", weights=None, unseen=True, generator=g)
    assert bp2 is None and bo2.shape == (90, 3)                                               # all 90 rays: a permutation
    assert np.array_equal(np.sort(bd2.cpu().numpy(), axis=0), np.sort(rows, axis=0))
