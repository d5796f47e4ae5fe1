"""Config 1 of SURVEY 8(d) (plumbing): the reference's training iteration (run_nerf_acc.py:284-307) on the CPU oracle path
makes the loss go down on an analytic phantom; on the GPU the same loop through the fused kernels does too, in both precision
modes, and ends close to the oracle's loss level."""
import functools

import numpy as np
import pytest
import torch

from oracle import cppn as ocppn, geometry as ogeo, nerfacc_ref, pipeline


def _phantom_targets(o, d, near, far, n=64):
    """Beer-Lambert projection of a ball of radius 40 (mu = 0.02) at the origin: analytic chord length."""
    b = np.einsum("ij,ij->i", o, d)
    dd = np.einsum("ij,ij->i", d, d)
    disc = b * b - dd * (np.einsum("ij,ij->i", o, o) - 40.0 ** 2)
    chord = np.where(disc > 0, 2 * np.sqrt(np.maximum(disc, 0)) / dd, 0.0)      # in units of t
    return np.exp(-0.02 * chord).astype(np.float32)


def _scene(W=12, views=((0.0, 0.0), (60.0, 0.0), (120.0, 0.0), (135.0, 135.0))):
    os_, ds_ = [], []
    for th, ph in views:
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], W, W, 7.5 * W)
        os_.append(o.reshape(-1, 3)); ds_.append(d.reshape(-1, 3))
    o = np.concatenate(os_).astype(np.float32); d = np.concatenate(ds_).astype(np.float32)
    return o, d, _phantom_targets(o.astype(np.float64), d.astype(np.float64), 1400.0, 1600.0)


def test_oracle_training_loss_decreases():
    o, d, target = _scene()
    roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
    grid = nerfacc_ref.OccupancyGrid(roi, 16)
    grid.binary[:] = True; grid.occs[:] = 0.05
    p = ocppn.init_params(2, 64, "none", 5, 0.05, seed=0)                       # 'small' CPPN, pos_enc none (config 1)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] - 4.0
    params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f = functools.partial(ocppn.cppn_forward, params, pos_enc="none", basis=5)
    opt = torch.optim.Adam(list(params.values()), lr=2e-3)
    tgt = torch.from_numpy(target)
    losses = []
    rng = np.random.default_rng(0)
    for it in range(40):
        sel = rng.permutation(len(o))[:256]
        pix, _ = pipeline.render_rays(f, grid, roi, o[sel], d[sel], 100, 1400.0, 1600.0, 1e-2, 1e-4)
        loss = torch.nn.functional.mse_loss(pix, tgt[sel])
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(float(loss))
    assert np.mean(losses[-5:]) < 0.6 * np.mean(losses[:5]), (losses[:5], losses[-5:])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gpu_training_loss_decreases(precision):
    import bench
    import nerf_for_angiography_b200 as A
    from nerf_for_angiography_b200.data import make_dataset
    from nerf_for_angiography_b200.train import Trainer
    dev = torch.device("cuda", 0)
    w = dict(bench.WORKLOADS["tiny"], rays=1024)
    pool, info = make_dataset(img_size=32, thetas=w["thetas"], kind="ct", volume_res=32, device=dev)
    torch.manual_seed(0)
    tr = Trainer(A.CPPN(bench.model_def(w, dev, precision)).to(dev), pool, info["near"], info["far"], n_rays=w["rays"], lr=5e-4, seed=0)
    losses = [float(tr.step()["loss"]) for _ in range(150)]
    assert np.isfinite(losses).all()
    assert np.mean(losses[-10:]) < 0.5 * np.mean(losses[:10]), (losses[:3], losses[-3:])
    if precision == "bf16":                      # attach the reference's weight image (mask -> EDT) for the vessel-pixel PSNR
        from nerf_for_angiography_b200.data import get_weighted_img
        pool.weights = torch.stack([torch.from_numpy(get_weighted_img(p.cpu().numpy())).float() for p in pool.pixels]).to(dev)
    ev = tr.evaluate()
    assert ev["psnr"] > 12.0 and ev["image"].shape == (32, 32)
    if precision == "bf16":
        assert ev["vessel_psnr"] is not None and np.isfinite(ev["vessel_psnr"])       # pixels with weight > the view's mean weight
    else:
        assert ev["vessel_psnr"] is None
