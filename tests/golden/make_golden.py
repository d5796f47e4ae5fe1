"""Generate tests/golden/*.npz by running the REFERENCE's own Python code (imported unmodified
from /root/reference) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the vectors are committed.  Third-party modules
that are absent here are shimmed exactly as SURVEY.md section 8c describes: ``matplotlib`` /
``skimage`` / ``nerfacc`` stubs (imported by the reference files but not used by the functions
called below) and ``torch_scatter.scatter_mul`` -> ``Tensor.scatter_reduce_('prod')``.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_shims():
    for name in ["matplotlib", "matplotlib.pyplot", "nerfacc", "skimage", "skimage.filters", "pyvista"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].filters = sys.modules["skimage.filters"]
    sys.modules["skimage.filters"].frangi = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("shim"))
    ts = types.ModuleType("torch_scatter")

    def scatter_mul(src, index, dim=0, out=None):
        return out.scatter_reduce_(dim, index, src, "prod", include_self=True)

    ts.scatter_mul = scatter_mul
    sys.modules["torch_scatter"] = ts


def main():
    _install_shims()
    sys.path.insert(0, os.path.join(REF, "nerf"))
    sys.path.insert(0, os.path.join(REF, "phantomdata"))
    sys.path.insert(0, REF)
    from model.CPPN import CPPN                       # /root/reference/model/CPPN.py
    import nerf_helpers                               # /root/reference/nerf/nerf_helpers.py
    import nerf_helpers_acc                           # /root/reference/nerf/nerf_helpers_acc.py
    import helpers as phantom_helpers                 # /root/reference/phantomdata/helpers.py
    import proj_helpers                               # /root/reference/phantomdata/proj_helpers.py

    dev = torch.device("cpu")

    # ---------------------------------------------------------------- geometry (a1, a2)
    geo = {}
    write_geometry = not any(a.startswith("--only=") for a in sys.argv[1:])
    cases = [(0.0, 0.0, 0.0, 16, 12, 120.0, (0, 0, 0)),
             (45.0, 0.0, 0.0, 12, 16, 90.0, (0, 0, 0)),
             (135.0, 135.0, 0.0, 10, 10, 75.0, (0, 0, 0)),
             (112.5, 22.5, 30.0, 9, 7, 1300.0, (1.5, -2.0, 3.0)),
             (354.0, 0.0, 0.0, 8, 8, 60.0, (0, 0, 0))]
    for ci, (th, ph, la, W, H, f, tr) in enumerate(cases):
        src = np.array([0.0, 0.0, 1500.0])
        o, d, M, ii, jj = phantom_helpers.get_ray_values(th, ph, la, src, W, H, f, dev, np.array(tr, dtype=np.float64))
        geo[f"c{ci}_args"] = np.array([th, ph, la, W, H, f, *tr], dtype=np.float64)
        geo[f"c{ci}_M"] = np.asarray(M, dtype=np.float64)
        geo[f"c{ci}_o"] = o.numpy().astype(np.float64)
        geo[f"c{ci}_d"] = d.numpy().astype(np.float64)
    geo["n_cases"] = np.array(len(cases))
    if write_geometry:
        np.savez_compressed(os.path.join(OUT, "geometry.npz"), **geo)

    # ---------------------------------------------------------------- CPPN (a9) fwd + grads
    only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]      # e.g. --only=cppn_fourier_2x256
    for tag, pos_enc, L, Hd in [("none_2x64", "none", 2, 64), ("fourier_4x128", "fourier", 4, 128),
                                ("fourier_2x64", "fourier", 2, 64), ("fourier_2x256", "fourier", 2, 256)]:
        if only and f"cppn_{tag}" not in only:
            continue
        torch.manual_seed(0)
        params = {'num_early_layers': L, 'num_late_layers': 0, 'num_filters': Hd, 'num_input_channels': 3,
                  'num_output_channels': 1, 'num_input_channels_views': 0, 'use_bias': True, 'pos_enc': pos_enc,
                  'pos_enc_basis': 5, 'act_func': 'relu', 'fourier_sigma': 5, 'num_img': 1, 'device': dev}
        model = CPPN(params)
        x = (torch.rand(192, 3) * 2 - 1) * 100.0
        y = model(x)
        # get_predictions with a ragged chunking must equal one-shot forward (a8)
        y_chunked = nerf_helpers.get_predictions(model, x, 50)
        assert torch.allclose(y, y_chunked, rtol=1e-5, atol=1e-6)   # CPU sgemm blocking differs per chunk
        g = torch.randn_like(y)
        (y * g).sum().backward()
        out = {"x": x.numpy(), "y": y.detach().numpy(), "y_chunked": y_chunked.detach().numpy(), "gout": g.numpy()}
        for k, v in model.state_dict().items():
            out["sd:" + k] = v.detach().numpy()
        for k, v in model.named_parameters():
            if v.grad is not None:
                out["grad:" + k] = v.grad.numpy()
        np.savez_compressed(os.path.join(OUT, f"cppn_{tag}.npz"), **out)
    if only:
        return

    # ---------------------------------------------------------------- CPPN with BARF encoding (f4): forward + grads at several alphas
    torch.manual_seed(0)
    params = {'num_early_layers': 4, 'num_late_layers': 0, 'num_filters': 128, 'num_input_channels': 3,
              'num_output_channels': 1, 'num_input_channels_views': 0, 'use_bias': True, 'pos_enc': 'barf',
              'pos_enc_basis': 5, 'act_func': 'relu', 'fourier_sigma': 5, 'num_img': 1, 'device': dev}
    model = CPPN(params)
    x = (torch.rand(160, 3) * 2 - 1) * 1.5           # BARF frequencies reach 16 pi: keep |x| small so fp32 phases stay meaningful
    out = {"x": x.numpy(), "alphas": np.array([0.0, 0.6, 1.0, 2.35, 3.999, 5.0, 7.0])}
    for k, v in model.state_dict().items():
        if k != "barf_weights":
            out["sd:" + k] = v.detach().numpy()
    for ai, a in enumerate(out["alphas"]):
        model.update_barf_alpha(float(a), 'pts')
        model.zero_grad()
        y = model(x)
        g = torch.randn(y.shape, generator=torch.Generator().manual_seed(100 + ai))
        (y * g).sum().backward()
        out[f"a{ai}_weights"] = model.barf_weights.detach().numpy()
        out[f"a{ai}_y"] = y.detach().numpy()
        out[f"a{ai}_gout"] = g.numpy()
        for k, v in model.named_parameters():
            if v.grad is not None and k != "barf_weights":
                out[f"a{ai}_grad:" + k] = v.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "cppn_barf_4x128.npz"), **out)

    # ---------------------------------------------------------------- composite (a10) + midpoints (a7)
    torch.manual_seed(1)
    n_rays = 37
    counts = torch.randint(0, 9, (n_rays,))
    counts[5] = 0
    counts[36] = 0
    ray_indices = torch.repeat_interleave(torch.arange(n_rays), counts)
    n = len(ray_indices)
    t_starts = 1400 + torch.rand(n, 1) * 190
    t_ends = t_starts + torch.rand(n, 1) * 2.0 + 0.1
    pred = torch.randn(n, 1, requires_grad=True)
    pix, ent = nerf_helpers_acc.acc_render_volume_density(pred, ray_indices, t_starts, t_ends, n_rays, 300)
    assert ent is None
    gp = torch.randn(n_rays)
    (pix * gp).sum().backward()
    with torch.no_grad():
        sig = torch.sigmoid(pred)
        zero_idx = torch.where(sig < 0.4)
        pix_zero, _ = nerf_helpers_acc.acc_render_volume_density(pred.detach().clone(), ray_indices, t_starts, t_ends,
                                                                 n_rays, 300, zero_idx)
    o = torch.randn(n_rays, 3) * 10 + torch.tensor([0.0, 0.0, 1500.0])
    d = torch.randn(n_rays, 3)
    positions = o[ray_indices.long()] + d[ray_indices.long()] * (t_starts + t_ends) / 2.0   # run_nerf_acc.py:290-292
    np.savez_compressed(os.path.join(OUT, "composite.npz"), ray_indices=ray_indices.numpy(), t_starts=t_starts.numpy(),
                        t_ends=t_ends.numpy(), pred=pred.detach().numpy(), pix=pix.detach().numpy(), gpix=gp.numpy(),
                        gpred=pred.grad.numpy(), pix_zero=pix_zero.numpy(), zero_mask=(sig < 0.4).numpy(),
                        n_rays=np.array(n_rays), o=o.numpy(), d=d.numpy(), positions=positions.numpy())
    # ---------------------------------------------------------------- phantom helpers (f2): transfer function + weight image
    rng = np.random.default_rng(7)
    hu = np.concatenate([rng.uniform(-500.0, 4500.0, 4000), np.array([0.0, 753.0, 1585.85, 2332.9, 3306.18, 4000.0, -1.0, 752.999, 3999.999])])
    tf = phantom_helpers.transfer_func_ct(hu)
    tf_bin = phantom_helpers.transfer_func_ct(hu, binary=True)
    sys.modules["matplotlib.pyplot"].imsave = lambda *a, **k: None          # get_weighted_img also writes a png
    img = np.ones((48, 40))
    yy, xx = np.mgrid[0:48, 0:40]
    img[(np.abs(xx - 12 - 0.3 * yy) < 2.5) | ((xx - 28) ** 2 + (yy - 30) ** 2 < 30)] = 0.35      # a vessel and a blob, pixel value < 1
    img += rng.uniform(-0.05, 0.0, img.shape) * (img < 1)
    wimg = phantom_helpers.get_weighted_img(img.copy(), 0.5, 0.5, 0, 0, 0, "/tmp/", sampling_strategy="segmentation")
    np.savez_compressed(os.path.join(OUT, "phantom.npz"), hu=hu, tf=tf, tf_bin=tf_bin, img=img, wimg=wimg)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
