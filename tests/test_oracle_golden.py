"""The oracle (CPU restatement) against golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import cppn, geometry, nerfacc_ref, pipeline


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_geometry_bit_exact(golden_dir):
    g = _load(golden_dir, "geometry.npz")
    for ci in range(int(g["n_cases"])):
        th, ph, la, W, H, f, tx, ty, tz = g[f"c{ci}_args"]
        o, d, M = geometry.get_ray_values(th, ph, la, np.array([0.0, 0.0, 1500.0]), int(W), int(H), f, (tx, ty, tz))
        assert np.array_equal(M, g[f"c{ci}_M"])
        assert o.shape == (int(H), int(W), 3)
        assert np.array_equal(o, g[f"c{ci}_o"])
        assert np.array_equal(d, g[f"c{ci}_d"])          # float64, bit for bit


@pytest.mark.parametrize("tag,pos_enc", [("none_2x64", "none"), ("fourier_4x128", "fourier"), ("fourier_2x64", "fourier"), ("fourier_2x256", "fourier")])
def test_cppn_forward_and_grads(golden_dir, tag, pos_enc):
    g = _load(golden_dir, f"cppn_{tag}.npz")
    params = {k[3:]: torch.from_numpy(g[k]).clone().requires_grad_(True) for k in g.files if k.startswith("sd:")}
    x = torch.from_numpy(g["x"])
    y = cppn.cppn_forward(params, x, pos_enc, 5)
    assert torch.equal(y.detach(), torch.from_numpy(g["y"]))   # same ops, same library => bit-exact on CPU
    (y * torch.from_numpy(g["gout"])).sum().backward()
    for k in g.files:
        if k.startswith("grad:"):
            ref = torch.from_numpy(g[k])
            got = params[k[5:]].grad
            assert got is not None
            assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6), k


def test_cppn_barf_forward_and_grads(golden_dir):
    """BARF encoding (f4): mask weights with the reference's quirks, forward and gradients at seven alphas, against the
    reference's own CPPN(pos_enc='barf')."""
    g = _load(golden_dir, "cppn_barf_4x128.npz")
    x = torch.from_numpy(g["x"])
    for ai, a in enumerate(g["alphas"]):
        w = cppn.barf_weights(float(a), 5)
        assert torch.equal(w, torch.from_numpy(g[f"a{ai}_weights"])), a
        params = {k[3:]: torch.from_numpy(g[k]).clone().requires_grad_(True) for k in g.files if k.startswith("sd:")}
        params["barf_weights"] = w
        y = cppn.cppn_forward(params, x, "barf", 5)
        assert torch.equal(y.detach(), torch.from_numpy(g[f"a{ai}_y"]))
        (y * torch.from_numpy(g[f"a{ai}_gout"])).sum().backward()
        for k in g.files:
            if k.startswith(f"a{ai}_grad:"):
                assert torch.allclose(params[k.split(":", 1)[1]].grad, torch.from_numpy(g[k]), rtol=1e-5, atol=1e-6), (a, k)


def test_composite_and_midpoints(golden_dir):
    g = _load(golden_dir, "composite.npz")
    n_rays = int(g["n_rays"])
    pred = torch.from_numpy(g["pred"]).clone().requires_grad_(True)
    ts, te = torch.from_numpy(g["t_starts"]), torch.from_numpy(g["t_ends"])
    pix = pipeline.acc_render_volume_density(pred, g["ray_indices"], ts, te, n_rays)
    assert torch.allclose(pix, torch.from_numpy(g["pix"]), rtol=1e-6, atol=1e-7)
    (pix * torch.from_numpy(g["gpix"])).sum().backward()
    assert torch.allclose(pred.grad, torch.from_numpy(g["gpred"]), rtol=1e-5, atol=1e-7)
    # rays without samples render exactly 1
    assert float(pix[5]) == 1.0 and float(pix[36]) == 1.0
    pz = pipeline.acc_render_volume_density(pred.detach(), g["ray_indices"], ts, te, n_rays,
                                            zero_mask=torch.from_numpy(g["zero_mask"]))
    assert torch.allclose(pz, torch.from_numpy(g["pix_zero"]), rtol=1e-6, atol=1e-7)
    # sequential-order scatter product (C) agrees with the library-order product
    alphas = torch.exp(-torch.sigmoid(pred.detach()) * (te - ts)).numpy()
    seq = nerfacc_ref.scatter_mul(alphas, g["ray_indices"], n_rays)
    assert np.allclose(seq, g["pix"], rtol=1e-6, atol=1e-7)
    pos = pipeline.midpoints(torch.from_numpy(g["o"]), torch.from_numpy(g["d"]), g["ray_indices"], ts, te)
    assert torch.equal(pos, torch.from_numpy(g["positions"]))
