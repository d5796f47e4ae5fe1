"""GPU parity tests of the composed hot path (reference call surface) against the CPU oracle:
CPPN module drop-in, bf16 tensor-core MLP vs fp32, render_rays, one full training step."""
import functools
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cppn as ocppn, geometry as ogeo, nerfacc_ref, pipeline  # noqa: E402


@pytest.fixture(scope="module")
def A():
    import nerf_for_angiography_b200 as a
    assert torch.cuda.is_available()
    return a


def _model_def(L, H, pos_enc, precision=None):
    d = {'num_early_layers': L, 'num_late_layers': 0, 'num_filters': H, 'num_input_channels': 3, 'num_output_channels': 1,
         'num_input_channels_views': 0, 'use_bias': True, 'pos_enc': pos_enc, 'pos_enc_basis': 5, 'act_func': 'relu',
         'fourier_sigma': 5, 'num_img': 1, 'device': torch.device("cuda")}
    if precision:
        d['precision'] = precision
    return d


def _load_golden_model(A, golden_dir, tag, L, H, pos_enc, precision):
    g = np.load(os.path.join(golden_dir, f"cppn_{tag}.npz"))
    model = A.CPPN(_model_def(L, H, pos_enc, precision))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    assert set(model.state_dict().keys()) == set(sd.keys())          # reference checkpoint keys, verbatim
    model.load_state_dict(sd)
    return model.to("cuda"), g


# ------------------------------------------------------------------------------------------------ CPPN drop-in
def test_cppn_module_matches_reference(A, golden_dir, tmp_path):
    model, g = _load_golden_model(A, golden_dir, "fourier_4x128", 4, 128, "fourier", "fp32")
    assert model.precision == "fp32"
    x = torch.from_numpy(g["x"]).cuda()
    y = model(x)
    assert y.shape == (x.shape[0], 1)
    ref = g["y"]
    assert np.max(np.abs(y.detach().cpu().numpy() - ref)) <= 1e-5 * max(1.0, np.abs(ref).max())
    (y * torch.from_numpy(g["gout"]).cuda()).sum().backward()
    for name, p in model.named_parameters():
        if "grad:" + name in g.files:
            gr = g["grad:" + name]
            assert p.grad is not None, name
            assert np.max(np.abs(p.grad.cpu().numpy() - gr)) <= 2e-5 * max(np.abs(gr).max(), 1e-6), name
    assert model.img1.grad is None                                    # dead parameters stay untouched (CPPN.py:134-135)
    # get_predictions: ragged chunking gives the same values (nerf_helpers.py:31-45)
    with torch.no_grad():
        yc = A.get_predictions(model, x, 50)
    assert torch.equal(yc, y.detach())
    # save() writes the reference's checkpoint dictionary (CPPN.py:261-276)
    f = str(tmp_path / "m.pth")
    model.save(f, {"note": 1})
    ck = torch.load(f, weights_only=False)
    assert set(ck.keys()) == {"version", "parameters", "training_information", "model"}
    m2 = A.CPPN(ck["parameters"]); m2.load_state_dict(ck["model"]); m2 = m2.to("cuda")
    with torch.no_grad():
        assert torch.equal(m2(x), y.detach())


def test_cppn_rejects_unsupported_configurations(A):
    d = _model_def(4, 128, "none"); d['act_func'] = 'sine'; d['sine_weights'] = 15
    with pytest.raises(NotImplementedError):
        A.CPPN(d)
    d = _model_def(4, 128, "none"); d['num_late_layers'] = 2         # skip connection
    with pytest.raises(NotImplementedError):
        A.CPPN(d)
    m = A.CPPN(_model_def(2, 64, "none")).to("cuda")
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 3))                                          # CPU tensor: no fallback


# ------------------------------------------------------------------------------------------------ bf16 tcgen05 forward
@pytest.mark.parametrize("pos_enc,n", [("fourier", 192), ("fourier", 128 * 148 * 2 * 3 + 77), ("none", 5000), ("fourier", 1)])
def test_mlp_bf16_tensor_core_forward(A, pos_enc, n):
    p = ocppn.init_params(4, 128, pos_enc, 5, 5.0, seed=1)
    model = A.CPPN(_model_def(4, 128, pos_enc, "bf16"))
    model.load_state_dict({**{k: v for k, v in p.items()}, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda")
    assert model.precision == "bf16"
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * 100.0
    with torch.no_grad():
        ref = ocppn.cppn_forward(p, x, pos_enc, 5).reshape(-1).numpy()
        y = model(x.cuda()).reshape(-1).cpu().numpy()
        sig = model.query(A.ops.OUT_SIGMA, points=x.cuda().contiguous()).cpu().numpy()
    scale = max(np.abs(ref).max(), 1e-3)
    err = np.abs(y - ref).max() / scale
    rms = np.sqrt(np.mean((y - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
    # bf16 operands, fp32 accumulate, six chained layers: worst element <= 4e-2 of the output scale, RMS <= 2e-2
    assert err <= 4e-2 and rms <= 2e-2, (err, rms)
    assert np.abs(sig - 1 / (1 + np.exp(-ref))).max() <= 1e-2


def test_mlp_bf16_ray_sample_inputs_and_alpha(A):
    p = ocppn.init_params(4, 128, "fourier", 5, 5.0, seed=2)
    model = A.CPPN(_model_def(4, 128, "fourier", "bf16"))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda")
    rng = np.random.default_rng(0)
    R, n = 64, 10000
    o = (rng.normal(size=(R, 3)) * 5 + [0, 0, 1500]).astype(np.float32)
    d = (rng.normal(size=(R, 3)) * 0.05 + [0, 0, -1]).astype(np.float32)
    ri = np.sort(rng.integers(0, R, n)).astype(np.int32)
    t0 = (1400 + rng.random(n) * 199).astype(np.float32); t1 = (t0 + 2.0 / 3.0).astype(np.float32)
    pos = pipeline.midpoints(torch.from_numpy(o), torch.from_numpy(d), ri, torch.from_numpy(t0)[:, None], torch.from_numpy(t1)[:, None])
    with torch.no_grad():
        logit = ocppn.cppn_forward(p, pos, "fourier", 5).reshape(-1)
        alpha = (1 - torch.exp(-torch.sigmoid(logit) * torch.from_numpy(t1 - t0))).numpy()
    kw = dict(rays_o=torch.from_numpy(o).cuda(), rays_d=torch.from_numpy(d).cuda(), ray_idx=torch.from_numpy(ri).cuda(),
              t_starts=torch.from_numpy(t0).cuda(), t_ends=torch.from_numpy(t1).cuda())
    got = model.query(A.ops.OUT_ALPHA, **kw).cpu().numpy()
    assert np.abs(got - alpha).max() <= 1e-2
    got_l = model.query(A.ops.OUT_LOGIT, **kw).cpu().numpy()
    assert np.abs(got_l - logit.numpy()).max() <= 4e-2 * max(1.0, np.abs(logit.numpy()).max())
    # device-resident sample count: capacity-sized arrays, the kernel stops at *n_dev
    n_dev = torch.tensor([7001], dtype=torch.int32, device="cuda")
    got_d = model.query(A.ops.OUT_ALPHA, n_dev=n_dev, **kw)
    assert got_d[:7001].cpu().numpy().tobytes() == got[:7001].tobytes()


# ------------------------------------------------------------------------------------------------ render_rays / training step
def _small_scene(A, res=32, W=24, seed=0):
    """oracle grid (numpy) + identical GPU grid, a handful of views' rays"""
    rng = np.random.default_rng(seed)
    roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
    r = np.arange(res)
    X, Y, Z = np.meshgrid(r, r, r, indexing="ij")
    binary = ((X - res / 2) ** 2 + (Y - res / 2) ** 2 + (Z - res / 2) ** 2 < (res / 3) ** 2)
    og = nerfacc_ref.OccupancyGrid(roi, res); og.binary = binary; og.occs[:] = 0.02
    gg = A.OccupancyGrid(torch.tensor(roi), res, A.ContractionType.AABB).cuda()
    gg._binary = torch.from_numpy(binary).cuda(); gg.occs.fill_(0.02); gg.occs_mean_host = 0.02
    os_, ds_ = [], []
    for th, ph in ((0.0, 0.0), (60.0, 0.0), (135.0, 135.0)):
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], W, W, 7.5 * W)
        os_.append(o.reshape(-1, 3)); ds_.append(d.reshape(-1, 3))
    o = np.concatenate(os_).astype(np.float32); d = np.concatenate(ds_).astype(np.float32)
    sel = rng.permutation(len(o))[:1024]
    return og, gg, roi, o[sel], d[sel]


def _scaled_params(L, H, pos_enc, seed):
    """random init, output bias shifted so sigma ~ 0.02-0.1: rays keep tens of samples before early stop"""
    p = ocppn.init_params(L, H, pos_enc, 5, 0.05, seed=seed)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] - 3.0
    return p


@pytest.mark.parametrize("precision,L,H,pos_enc", [("fp32", 2, 64, "none"), ("fp32", 4, 128, "fourier"), ("bf16", 4, 128, "fourier")])
def test_render_rays_vs_oracle(A, precision, L, H, pos_enc):
    og, gg, roi, o, d = _small_scene(A)
    p = _scaled_params(L, H, pos_enc, 5)
    model = A.CPPN(_model_def(L, H, pos_enc, precision))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda")
    f = functools.partial(ocppn.cppn_forward, p, pos_enc=pos_enc, basis=5)
    with torch.no_grad():
        pix_ref, (ri, ts, te) = pipeline.render_rays(f, og, roi, o, d, 300, 1400.0, 1600.0, 1e-2, 1e-4)
    oc, dc = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    with torch.no_grad():
        pix, (gi, g0, g1) = A.render_rays(model, gg, torch.tensor(roi).cuda(), oc, dc, 300, 1400.0, 1600.0, 1e-2, 1e-4)
    assert gi.dtype == torch.int64 and g0.shape == (len(gi), 1)
    pix, pix_ref = pix.cpu().numpy(), pix_ref.numpy()
    if precision == "fp32":
        # sample set: identical up to visibility decisions that sit within rounding of a threshold -- the symmetric difference of
        # the (ray, t_start) sets is bounded, and every sample both sides kept has bit-identical interval ends
        key = lambda r, t: (np.asarray(r, np.int64) << 32) | np.asarray(t, np.float32).reshape(-1).view(np.uint32).astype(np.int64)
        a, b = key(gi.cpu().numpy(), g0.cpu().numpy()), key(ri, ts)
        assert np.setxor1d(a, b).size <= max(2, int(1e-4 * len(ri))), np.setxor1d(a, b).size
        _, ia, ib = np.intersect1d(a, b, return_indices=True)
        assert len(ia) >= len(ri) - max(2, int(1e-4 * len(ri)))
        assert np.array_equal(g1.cpu().numpy().reshape(-1)[ia], te.reshape(-1)[ib])
        assert np.max(np.abs(pix - pix_ref)) <= 1e-5                 # fp32 check mode: <= 1e-5 on the projection
    else:
        # bf16 MLP: <= 1e-2 relative error on the projection, measured against the image scale (pixels live in [0, 1])
        # and as a relative L2 norm; per-pixel ratios on nearly black pixels (value 0.01) only amplify 1e-4 absolute noise
        assert np.max(np.abs(pix - pix_ref)) <= 1e-2 * pix_ref.max()
        assert np.linalg.norm(pix - pix_ref) / np.linalg.norm(pix_ref) <= 1e-2
        bright = pix_ref > 0.1
        assert np.max(np.abs(pix - pix_ref)[bright] / pix_ref[bright]) <= 1e-2


def test_training_step_vs_oracle_fp32(A):
    """one reference iteration (march, filter, forward, composite, mse, backward, Adam) on fixed rays"""
    from nerf_for_angiography_b200.train import Trainer
    from nerf_for_angiography_b200.data import RayPool
    og, gg, roi, o, d = _small_scene(A, seed=3)
    L, H, pos_enc = 2, 64, "fourier"
    p = _scaled_params(L, H, pos_enc, 7)
    target = np.random.default_rng(1).random(len(o)).astype(np.float32)
    # oracle step (torch CPU autograd + torch.optim.Adam)
    params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f = functools.partial(ocppn.cppn_forward, params, pos_enc=pos_enc, basis=5)
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    pix_ref, (ri, ts, te) = pipeline.render_rays(f, og, roi, o, d, 300, 1400.0, 1600.0, 1e-2, 1e-4)
    loss_ref = torch.nn.functional.mse_loss(pix_ref, torch.from_numpy(target))
    opt.zero_grad(); loss_ref.backward(); opt.step()
    # B200 step
    model = A.CPPN(_model_def(L, H, pos_enc, "fp32"))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda")
    dummy_pool = RayPool(torch.eye(4, dtype=torch.float64).cuda()[None], torch.zeros(1, 2, 2).cuda(), 1.0)
    tr = Trainer(model, dummy_pool, 1400.0, 1600.0, n_rays=len(o), vessel_grid=False)
    tr.acc_grid = gg
    tr.n_iter = 1                                                      # skip the grid refresh (step % 16 != 0)
    out = tr.step(rays=(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(target).cuda()))
    assert abs(out["n_samples"] - len(ri)) <= 2
    assert np.isclose(float(out["loss"]), float(loss_ref), rtol=1e-5)
    # gradients (flat layout: coef, W0, b0, ..., W_out, b_out) before the optimiser touched them
    order = (["fourier_coefficients"] if pos_enc == "fourier" else []) + \
        [f"early_pts_layers.{2 * i}.{w}" for i in range(L + 1) for w in ("weight", "bias")] + ["output_linear.0.weight", "output_linear.0.bias"]
    gref = np.concatenate([params[k].grad.numpy().reshape(-1) for k in order])
    got = tr.grad[:-1].cpu().numpy()                                 # last slot: kept-sample count of the sync-free loop
    assert np.max(np.abs(got - gref)) <= 1e-4 * np.abs(gref).max(), np.max(np.abs(got - gref)) / np.abs(gref).max()
    # parameters after Adam: wherever the gradient is well above Adam's eps the update is -lr*sign(g) on both sides
    sd = model.state_dict()
    for k in order:
        g = params[k].grad.numpy()
        big = np.abs(g) > 1e-5
        assert np.allclose(sd[k].cpu().numpy()[big], params[k].detach().numpy()[big], rtol=0, atol=2e-6), k


def test_sync_free_step_equals_one_sync_step(A):
    """bf16 training step with every sample count kept on the device (capacity-sized buffers, n_dev everywhere) against the
    same step with the kept count read back: same samples, bit-identical loss, gradients and parameters.  Also the empty
    iteration: no samples -> no optimiser step."""
    from nerf_for_angiography_b200.data import RayPool
    from nerf_for_angiography_b200.train import Trainer
    og, gg, roi, o, d = _small_scene(A)
    p = _scaled_params(4, 128, "fourier", seed=5)
    target = torch.from_numpy(np.random.default_rng(2).random(len(o)).astype(np.float32)).cuda()
    rays = (torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), target)
    res = {}
    for mode in (False, True):
        model = A.CPPN(_model_def(4, 128, "fourier", "bf16"))
        model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
        model = model.to("cuda")
        pool = RayPool(torch.eye(4, dtype=torch.float64).cuda()[None], torch.zeros(1, 2, 2).cuda(), 1.0)
        tr = Trainer(model, pool, 1400.0, 1600.0, n_rays=len(o), vessel_grid=False, sync_free=mode)
        assert tr.sync_free is mode
        tr.acc_grid = gg
        tr.n_iter = 1
        out = tr.step(rays=rays)
        res[mode] = (out["n_samples_prefilter"], out["n_samples"], float(out["loss"]), tr.grad[:-1].clone(), tr.flat.clone(), out["pix"].clone())
    a, b = res[False], res[True]
    # the sync-free loop marches lazily: its pre-filter count is the samples that were marched at all (<= the full march)
    assert a[0] >= b[0] > 0 and a[1] == b[1] > 0
    assert a[5].equal(b[5]) and np.isclose(a[2], b[2], rtol=1e-6)     # same samples -> bit-identical projection (the loss sum uses float atomics)
    # gradients: the persistent kernels spread the tiles over a capacity-sized grid, so the fixed-order partial sums of the
    # reductions are grouped differently -> fp32 re-association only (tolerance 1e-5 of the largest gradient entry)
    scale = float(a[3].abs().max())
    assert float((a[3] - b[3]).abs().max()) <= 1e-5 * scale, float((a[3] - b[3]).abs().max()) / scale
    assert float((a[4] - b[4]).abs().max()) <= 2.1e-4            # one Adam step moves a parameter by at most lr = 1e-4
    # empty grid: no samples anywhere -> parameters and Adam moments untouched, pixels render 1
    empty = A.OccupancyGrid(torch.tensor(roi), 32, A.ContractionType.AABB).cuda()
    tr.acc_grid = empty
    before = tr.flat.clone()
    out = tr.step(rays=rays)
    assert out["n_samples"] == 0 and tr.flat.equal(before) and bool((out["pix"] == 1).all())


@pytest.mark.parametrize("bias_shift", [0.0, -3.0, -9.0])
def test_two_phase_visibility_equals_full_evaluation(A, bias_shift):
    """Early ray termination: alphas only for the first 32 samples of each ray + the rest of the rays still alive, against
    alphas for every sample -- same keep mask, same compacted samples, bit for bit (dense field: all rays die early;
    thin field: all survive; in between: mixed)."""
    og, gg, roi, o, d = _small_scene(A, res=32, W=40)
    p = ocppn.init_params(4, 128, "fourier", 5, 0.05, seed=7)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] + bias_shift
    model = A.CPPN(_model_def(4, 128, "fourier", "bf16"))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda"); model._ensure_flat()
    packed = A.ops.mlp_pack(model._desc, model._flat)
    ro, rd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    aabb = np.array(roi, np.float32)
    step = 200.0 / 300
    ri, t0, t1, off = A.ops.march(ro, rd, aabb, aabb, 32, gg._binary_u8(), 1400.0, 1600.0, step)
    kw = dict(rays_o=ro, rays_d=rd, ray_idx=ri, t_starts=t0, t_ends=t1)
    full = A.ops.mlp_forward(model._desc, model._flat, packed, A.ops.OUT_ALPHA, A.ops.PREC_BF16, **kw)
    two, evaluated = A.ops.alphas_two_phase(model._desc, model._flat, packed, A.ops.PREC_BF16, ro, rd, ri, t0, t1, off, 1e-2, k0=32)
    ev = evaluated.tolist()
    assert ev[0] == int(torch.clamp(off[1:] - off[:-1], max=32).sum()) and ev[0] + ev[1] <= ri.numel()
    a = A.ops.visibility_compact(full, off, t0, t1, 1e-2, 1e-4)
    b = A.ops.visibility_compact(two, off, t0, t1, 1e-2, 1e-4)
    for x, y in zip(a, b):
        assert x.equal(y)
    if bias_shift == 0.0:
        assert ev[1] < 0.5 * (ri.numel() - ev[0])             # dense field: most rays are opaque after 32 samples
    if bias_shift == -9.0:
        assert ev[0] + ev[1] == ri.numel()                    # thin field: nothing terminates, every sample is evaluated


# bf16: the golden inputs are 160 points near the origin (|logit| ~ 0.01, many units at the ReLU threshold), so bf16 rounding flips
# masks that do not average out over so few samples: outputs agree to 1e-4 absolute, per-tensor gradients to 3-14 % (tools/barf_diag.py);
# the 3e-2 bound of the training-size tests (test_mlp_bf16_tensor_core_backward) does not apply here.  fp32 is exact to 1e-4.
@pytest.mark.parametrize("precision,tol_y,tol_g", [("fp32", 2e-5, 1e-4), ("bf16", 2e-2, 2e-1)])
def test_cppn_barf_matches_reference_golden(A, golden_dir, precision, tol_y, tol_g):
    """BARF encoding (coarse-to-fine mask, /root/reference/model/CPPN.py:82-94,224-259) on the fused kernels against golden
    vectors produced by the reference's own CPPN(pos_enc='barf'): forward and parameter gradients at seven mask positions."""
    g = np.load(os.path.join(golden_dir, "cppn_barf_4x128.npz"))
    mdef = _model_def(4, 128, "barf", precision)
    model = A.CPPN(mdef)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    model.load_state_dict({**sd, "barf_weights": torch.zeros(15)})
    model = model.to("cuda")
    assert "barf_weights" in model.state_dict() and "fourier_coefficients" not in model.state_dict()
    x = torch.from_numpy(g["x"]).cuda()
    for ai, a in enumerate(g["alphas"]):
        model.update_barf_alpha(float(a), 'pts')
        assert torch.equal(model.barf_weights.cpu(), torch.from_numpy(g[f"a{ai}_weights"]))
        model.zero_grad()
        y = model(x)
        yref = torch.from_numpy(g[f"a{ai}_y"])
        assert float((y.detach().cpu() - yref).abs().max()) <= tol_y * max(1.0, float(yref.abs().max())), a
        (y * torch.from_numpy(g[f"a{ai}_gout"]).cuda()).sum().backward()
        for k in g.files:
            if k.startswith(f"a{ai}_grad:"):
                name = k.split(":", 1)[1]
                got = dict(model.named_parameters())[name].grad.cpu()
                ref = torch.from_numpy(g[k])
                assert float((got - ref).norm()) <= tol_g * max(float(ref.norm()), 1e-6), (a, name)


def test_trainer_with_barf_schedule(A):
    """The training loop with a BARF model: the mask is ramped like the reference driver does (run_nerf_acc.py:268-272), the
    fixed frequencies never move, the loss goes down."""
    import bench
    from nerf_for_angiography_b200.data import make_dataset
    from nerf_for_angiography_b200.train import Trainer
    dev = torch.device("cuda", 0)
    w = dict(bench.WORKLOADS["tiny"], rays=1024)
    pool, info = make_dataset(img_size=32, thetas=w["thetas"], kind="ct", volume_res=32, device=dev)
    torch.manual_seed(0)
    mdef = bench.model_def(w, dev, "bf16"); mdef['pos_enc'] = 'barf'
    model = A.CPPN(mdef).to(dev)
    tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"], lr=5e-4, seed=0)
    coef0 = tr.flat[:15].clone()
    assert torch.equal(coef0.cpu(), 2.0 ** (torch.repeat_interleave(torch.arange(0., 5), 3) - 1))
    losses = []
    for it in range(120):
        model.update_barf_alpha(6.0 * it / 100, 'pts')                # coarse to fine over the first 100 iterations (full at alpha = basis + 1)
        losses.append(float(tr.step()["loss"]))
    assert np.isfinite(losses).all() and np.mean(losses[-10:]) < 0.6 * np.mean(losses[:10])
    assert torch.equal(tr.flat[:15], coef0) and bool((model.barf_weights == 1).all())


def test_two_phase_visibility_ragged_rays(A):
    """Rays with fewer than 32 samples, exactly 32, and none at all (single occupied slab): the index lists stay consistent and
    the result still equals the full evaluation."""
    res = 32
    binary = np.zeros((res,) * 3, bool)
    binary[14:18, :, :] = True                                          # a 25-unit slab: 0 .. ~40 samples per ray
    roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
    gg = A.OccupancyGrid(torch.tensor(roi), res, A.ContractionType.AABB).cuda()
    gg._binary = torch.from_numpy(binary).cuda()
    os_, ds_ = [], []
    for th, ph in ((0.0, 0.0), (70.0, 20.0), (135.0, 135.0)):
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], 32, 32, 7.5 * 32)
        os_.append(o.reshape(-1, 3)); ds_.append(d.reshape(-1, 3))
    ro = torch.from_numpy(np.concatenate(os_).astype(np.float32)).cuda()
    rd = torch.from_numpy(np.concatenate(ds_).astype(np.float32)).cuda()
    p = ocppn.init_params(4, 128, "fourier", 5, 0.05, seed=3)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] - 2.0
    model = A.CPPN(_model_def(4, 128, "fourier", "bf16"))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda"); model._ensure_flat()
    packed = A.ops.mlp_pack(model._desc, model._flat)
    ri, t0, t1, off = A.ops.march(ro, rd, roi, roi, res, gg._binary_u8(), 1400.0, 1600.0, 200.0 / 300)
    cnt = (off[1:] - off[:-1])
    assert int(cnt.min()) == 0 and int((cnt < 32).sum()) > 100 and int((cnt > 32).sum()) > 100
    kw = dict(rays_o=ro, rays_d=rd, ray_idx=ri, t_starts=t0, t_ends=t1)
    full = A.ops.mlp_forward(model._desc, model._flat, packed, A.ops.OUT_ALPHA, A.ops.PREC_BF16, **kw)
    two, ev = A.ops.alphas_two_phase(model._desc, model._flat, packed, A.ops.PREC_BF16, ro, rd, ri, t0, t1, off, 1e-2, k0=32)
    assert ev.tolist()[0] == int(torch.clamp(cnt, max=32).sum())
    for x, y in zip(A.ops.visibility_compact(full, off, t0, t1, 1e-2, 1e-4), A.ops.visibility_compact(two, off, t0, t1, 1e-2, 1e-4)):
        assert x.equal(y)


def _slab_scene(A, res=32):
    binary = np.zeros((res,) * 3, bool)
    binary[14:18, :, :] = True                                          # a 25-unit slab: 0 .. ~40 samples per ray
    roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
    gg = A.OccupancyGrid(torch.tensor(roi), res, A.ContractionType.AABB).cuda()
    gg._binary = torch.from_numpy(binary).cuda()
    os_, ds_ = [], []
    for th, ph in ((0.0, 0.0), (70.0, 20.0), (135.0, 135.0)):
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], 32, 32, 7.5 * 32)
        os_.append(o.reshape(-1, 3)); ds_.append(d.reshape(-1, 3))
    return gg, roi, np.concatenate(os_).astype(np.float32), np.concatenate(ds_).astype(np.float32)


@pytest.mark.parametrize("scene,bias_shift,k0", [("dense", 0.0, 32), ("dense", -3.0, 32), ("dense", -9.0, 32), ("dense", -3.0, 8),
                                                 ("slab", -2.0, 32), ("slab", -9.0, 5), ("slab", 0.0, 1)])
def test_lazy_marching_equals_full_march_and_filter(A, scene, bias_shift, k0):
    """Lazy marching (head of k0 samples per ray -> alpha -> visibility -> tail of the rays still alive -> one compaction) against
    the reference order (march every ray to the end, alpha for every sample, filter, compact): the kept samples, their ray
    indices, the offsets and the kept count are bit-identical; the marched count is never larger."""
    if scene == "dense":
        og, gg, roi, o, d = _small_scene(A, res=32, W=40)
        res = 32
    else:
        gg, roi, o, d = _slab_scene(A)
        res = 32
    p = ocppn.init_params(4, 128, "fourier", 5, 0.05, seed=7)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] + bias_shift
    model = A.CPPN(_model_def(4, 128, "fourier", "bf16"))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda"); model._ensure_flat()
    packed = A.ops.mlp_pack(model._desc, model._flat)
    ro, rd = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    aabb = np.array(roi, np.float32)
    step = 200.0 / 300
    ri, t0, t1, off = A.ops.march(ro, rd, aabb, aabb, res, gg._binary_u8(), 1400.0, 1600.0, step)
    full = A.ops.mlp_forward(model._desc, model._flat, packed, A.ops.OUT_ALPHA, A.ops.PREC_BF16, rays_o=ro, rays_d=rd, ray_idx=ri,
                             t_starts=t0, t_ends=t1)
    e_ri, e_t0, e_t1, e_off, _ = A.ops.visibility_compact(full, off, t0, t1, 1e-2, 1e-4)
    totals = torch.zeros(4, dtype=torch.int32, device="cuda")
    l_ri, l_t0, l_t1, l_off = A.ops.march_filter_lazy(model._desc, model._flat, packed, A.ops.PREC_BF16, ro, rd, aabb, aabb, res,
                                                      gg._binary_u8(), 1400.0, 1600.0, step, 1e-2, 1e-4, k0=k0, totals=totals)
    n = e_ri.numel()
    marched, kept = totals.tolist()[:2]
    assert kept == n == int(l_off[-1]) and l_off.equal(e_off)
    assert l_ri[:n].equal(e_ri) and l_t0[:n].equal(e_t0) and l_t1[:n].equal(e_t1)
    cnt = off[1:] - off[:-1]
    assert int(torch.clamp(cnt, max=k0).sum()) <= marched <= ri.numel()
    if scene == "dense" and bias_shift == 0.0:
        assert marched < 0.6 * ri.numel()                     # dense field: most rays are opaque behind their head
    if bias_shift == -9.0:
        assert marched == ri.numel()                          # thin field: every ray is marched to its end


def test_checkpoint_resume_is_exact(A, tmp_path):
    """save_checkpoint -> fresh Trainer.load_checkpoint -> the continued run is bit-identical to the uninterrupted one; the
    file has the reference's .pth layout (model/CPPN.py:261-276) and its state_dict loads into a new CPPN."""
    import bench
    from nerf_for_angiography_b200.data import make_dataset
    from nerf_for_angiography_b200.train import Trainer
    dev = torch.device("cuda", 0)
    w = dict(bench.WORKLOADS["tiny"], rays=1024)
    pool, info = make_dataset(img_size=32, thetas=w["thetas"], kind="ct", volume_res=32, device=dev)

    def make():
        torch.manual_seed(0)
        return Trainer(A.CPPN(bench.model_def(w, dev, "bf16")).to(dev), pool, info["near"], info["far"], n_rays=w["rays"], seed=0)
    tr = make()
    for _ in range(18):                                                # crosses a grid refresh (iteration 16)
        tr.step()
    path = str(tmp_path / "ck.pth")
    tr.save_checkpoint(path, extra={"note": "unit test"})
    ref_losses = [float(tr.step()["loss"]) for _ in range(3)]
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"version", "parameters", "training_information", "model"} and ck["training_information"]["note"] == "unit test"
    assert {"early_pts_layers.0.weight", "output_linear.0.bias", "fourier_coefficients", "img1", "img2"} <= set(ck["model"])
    tr2 = make()
    pool._seed_streams = {}
    tr2.load_checkpoint(path)
    got = [float(tr2.step()["loss"]) for _ in range(3)]
    assert np.allclose(got, ref_losses, rtol=1e-6, atol=0)           # the scalar loss is summed with float atomics
    assert tr2.acc_grid.occs.equal(tr.acc_grid.occs) and tr2.n_iter == tr.n_iter
    assert tr2.flat.equal(tr.flat)                                     # same rays, same kernels, fixed-order reductions
    m = A.CPPN(ck["parameters"]); m.load_state_dict(ck["model"])      # what the reference does with a checkpoint


# ------------------------------------------------------------------------------------------------ bf16 tcgen05 backward
@pytest.mark.parametrize("pos_enc,L,n", [("fourier", 4, 128 * 148 * 2 + 300), ("none", 4, 5000), ("fourier", 2, 777), ("fourier", 4, 100)])
def test_mlp_bf16_tensor_core_backward(A, pos_enc, L, n):
    """tcgen05 forward(train) + dgrad + wgrad against the fp32 check path on the same parameters and samples"""
    p = ocppn.init_params(L, 128, pos_enc, 5, 0.2, seed=4)
    model = A.CPPN(_model_def(L, 128, pos_enc, "bf16"))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda")
    model._ensure_flat()
    desc, flat = model._desc, model._flat
    packed = A.ops.mlp_pack(desc, flat)
    rng = np.random.default_rng(n)
    R = 97
    o = (rng.normal(size=(R, 3)) * 5 + [0, 0, 1500]).astype(np.float32)
    d = (rng.normal(size=(R, 3)) * 0.05 + [0, 0, -1]).astype(np.float32)
    ri = np.sort(rng.integers(0, R, n)).astype(np.int32)
    t0 = (1400 + rng.random(n) * 199).astype(np.float32); t1 = (t0 + 2.0 / 3.0).astype(np.float32)
    kw = dict(rays_o=torch.from_numpy(o).cuda(), rays_d=torch.from_numpy(d).cuda(), ray_idx=torch.from_numpy(ri).cuda(),
              t_starts=torch.from_numpy(t0).cuda(), t_ends=torch.from_numpy(t1).cuda())
    g = torch.from_numpy(rng.normal(size=n).astype(np.float32)).cuda()
    y32, s32 = A.ops.mlp_forward(desc, flat, None, A.ops.OUT_LOGIT, A.ops.PREC_FP32, saved=True, **kw)
    g32 = A.ops.mlp_backward(desc, flat, None, s32, g, A.ops.PREC_FP32, **kw)
    y16, s16 = A.ops.mlp_forward(desc, flat, packed, A.ops.OUT_LOGIT, A.ops.PREC_BF16, saved=True, **kw)
    y16i = A.ops.mlp_forward(desc, flat, packed, A.ops.OUT_LOGIT, A.ops.PREC_BF16, **kw)
    assert torch.equal(y16, y16i)                                      # training forward == inference forward, bit for bit
    assert float((y16 - y32).abs().max()) <= 4e-2 * max(1.0, float(y32.abs().max()))
    g16 = A.ops.mlp_backward(desc, flat, packed, s16, g, A.ops.PREC_BF16, **kw)
    # (a) against the fp32 path end to end.  With a random upstream gradient every weight gradient is a random-walk sum,
    # so the ~1-2 % of ReLU units whose sign differs between the bf16 and fp32 forward move it by sqrt(1-2 %) ~ 10 %:
    # this bounds gross errors only.
    names = [k for k, _ in model.named_parameters() if not k.startswith("img")]
    rel_e2e = {nm: float((g16[o_:o_ + c_] - g32[o_:o_ + c_]).norm() / g32[o_:o_ + c_].norm().clamp_min(1e-12))
               for (o_, c_, _), nm in zip(model._param_slices, names)}
    assert max(rel_e2e.values()) <= 0.25, rel_e2e
    # (b) the backward kernels themselves: run the fp32 backward on the SAME (bf16) activations the tensor-core forward
    # saved (tile images decoded here), so both sides see identical ReLU masks: <= 2e-2 relative L2 per tensor.
    s32b = _decode_saved_images(s16, n, L, 5 if pos_enc == "fourier" else 0)
    g32b = A.ops.mlp_backward(desc, flat, None, s32b, g, A.ops.PREC_FP32, **kw)
    rels = {nm: round(float((g16[o_:o_ + c_] - g32b[o_:o_ + c_]).norm() / g32b[o_:o_ + c_].norm().clamp_min(1e-12)), 4)
            for (o_, c_, _), nm in zip(model._param_slices, names)}
    print(rels)
    assert max(rels.values()) <= 2e-2, rels


def _decode_saved_images(saved_u8, n, L, basis):
    """bf16 swizzled tile images (csrc/mlp_tc.cu) -> the fp32 path's saved layout [X0 | a_1 | ... | a_{L+1}]"""
    n_tiles = (n + 127) // 128
    raw = saved_u8.cpu().numpy()

    def block(off, rows_tiles):                      # [n_tiles*128, 64] float32 from consecutive 16 KB blocks with given byte offsets
        out = np.empty((n_tiles * 128, 64), np.float32)
        r = np.arange(128)[:, None]; c = np.arange(64)[None, :]
        byte = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2
        for t in range(n_tiles):
            b = raw[rows_tiles(t):rows_tiles(t) + 16384]
            u16 = b.view(np.uint16)[(byte // 2)]
            out[t * 128:(t + 1) * 128] = (u16.astype(np.uint32) << 16).view(np.float32)
        return out

    a0 = block(0, lambda t: t * 16384)
    d_in = 3 + 6 * basis
    X0 = np.zeros((n, d_in), np.float32)
    X0[:, :3] = (a0[:n, 0:3] + a0[:n, 3:6])
    nb = 3 * basis
    for j in range(nb):
        X0[:, 3 + j] = a0[:n, 6 + 2 * j]
        X0[:, 3 + nb + j] = a0[:n, 7 + 2 * j]
    al = lambda v: (v + 255) // 256 * 256          # noqa: E731
    parts = [np.zeros(al(n * d_in * 4), np.uint8)]
    parts[0][:n * d_in * 4] = X0.reshape(-1).view(np.uint8)
    base = n_tiles * 16384
    for l in range(L + 1):
        lo = block(0, lambda t, l=l: base + (l * n_tiles + t) * 32768)
        hi = block(0, lambda t, l=l: base + (l * n_tiles + t) * 32768 + 16384)
        act = np.concatenate([lo, hi], axis=1)[:n]
        buf = np.zeros(al(n * 128 * 4), np.uint8)
        buf[:n * 128 * 4] = np.ascontiguousarray(act).reshape(-1).view(np.uint8)
        parts.append(buf)
    return torch.from_numpy(np.concatenate(parts)).cuda()


# ------------------------------------------------------------------------------------------------ inference entry points
def test_inference_render_projections_and_volume_query(A):
    og, gg, roi, _, _ = _small_scene(A)
    p = _scaled_params(2, 64, "fourier", 9)
    model = A.CPPN(_model_def(2, 64, "fourier", "fp32"))
    model.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    model = model.to("cuda")
    f = functools.partial(ocppn.cppn_forward, p, pos_enc="fourier", basis=5)
    W, H, focal = 14, 10, 105.0                                        # W != H: catches x/y transposition
    views = [(30.0, 0.0), (135.0, 135.0)]
    imgs, bins = A.render_projections(model, gg, torch.tensor(roi).cuda(), views, np.array([0, 0, 1500.0]), W, H, focal, 300, 1400.0,
                                      1600.0, 1e-2, 1e-4, binary_thresh=0.05)
    assert imgs.shape == (2, H, W)
    for v, (th, ph) in enumerate(views):
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], W, H, focal)
        with torch.no_grad():
            ref, (ri, ts, te) = pipeline.render_rays(f, og, roi, o.reshape(-1, 3).astype(np.float32), d.reshape(-1, 3).astype(np.float32),
                                                     300, 1400.0, 1600.0, 1e-2, 1e-4)
        assert np.max(np.abs(imgs[v].cpu().numpy().reshape(-1) - ref.numpy())) <= 1e-5
        assert float(bins[v].min()) >= float(imgs[v].min()) - 1e-6     # zeroing sigma can only brighten the projection
    # volume query on the reference's 'xy' meshgrid lattice
    t = np.linspace(-75, 75, 7).astype(np.float32)
    vol = A.query_volume(model, t).cpu().numpy()
    mesh = np.stack(np.meshgrid(t, t, t), -1).reshape(-1, 3)           # visualization.py:209-211
    with torch.no_grad():
        ref = torch.sigmoid(f(torch.from_numpy(mesh))).reshape(7, 7, 7).numpy()
    assert np.max(np.abs(vol - ref)) <= 1e-5
    masked = A.query_volume(model, t, grid=gg).cpu().numpy()
    occ = og.query_occ(mesh).reshape(7, 7, 7)
    assert np.max(np.abs(masked - ref * occ)) <= 1e-5
