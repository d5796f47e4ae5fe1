"""GPU parity tests of the width-256 tcgen05 kernels (BASELINE config 4: 8 x 256 CPPN, csrc/mlp_tc256.cu) against the CPU oracle
and the fp32 check path: forward (points / ray samples / every output mode / device-resident counts / index lists), training
forward == inference forward bit for bit, data- and weight-gradient kernels on identical ReLU masks, render_rays, and one full
training step.  Reference: /root/reference/model/CPPN.py:96-131,166-222 (the model is width / depth generic)."""
import functools

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


def _mdef(L, pos_enc, precision):
    return {'num_early_layers': L, 'num_late_layers': 0, 'num_filters': 256, 'num_input_channels': 3, 'num_output_channels': 1,
            'num_input_channels_views': 0, 'use_bias': True, 'pos_enc': pos_enc, 'pos_enc_basis': 5, 'act_func': 'relu',
            'fourier_sigma': 5, 'num_img': 1, 'device': torch.device("cuda"), 'precision': precision}


def _model(A, p, L, pos_enc, precision):
    m = A.CPPN(_mdef(L, pos_enc, precision))
    m.load_state_dict({**p, "img1": torch.zeros(2), "img2": torch.zeros(2)})
    return m.to("cuda")


def test_width256_is_on_the_tensor_core_path(A):
    m = A.CPPN(_mdef(8, "fourier", None)).to("cuda")
    assert m.precision == "bf16"                                   # default precision: tcgen05 path, not the fp32 check path
    assert A.ops.mlp_param_count(m._desc) == 535316 - 4            # reference count minus the dead img1 / img2


@pytest.mark.parametrize("L,pos_enc,n", [(8, "fourier", 128 * 148 * 2 + 77), (8, "fourier", 1), (2, "none", 5000), (1, "fourier", 300)])
def test_mlp256_forward_vs_oracle(A, L, pos_enc, n):
    p = ocppn.init_params(L, 256, pos_enc, 5, 5.0, seed=1)
    model = _model(A, p, L, pos_enc, "bf16")
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * 100.0
    with torch.no_grad():
        ref = ocppn.cppn_forward(p, x, pos_enc, 5).reshape(-1).numpy()
        y = model(x.cuda()).reshape(-1).cpu().numpy()
        sig = model.query(A.ops.OUT_SIGMA, points=x.cuda().contiguous()).cpu().numpy()
    scale = max(np.abs(ref).max(), 1e-3)
    err = np.abs(y - ref).max() / scale
    rms = np.sqrt(np.mean((y - ref) ** 2)) / max(np.sqrt(np.mean(ref ** 2)), 1e-6)
    # bf16 operands, fp32 accumulate, up to ten chained layers of K = 256: worst element <= 5e-2 of the output scale, RMS <= 2.5e-2
    assert err <= 5e-2 and rms <= 2.5e-2, (err, rms)
    assert np.abs(sig - 1 / (1 + np.exp(-ref))).max() <= 1.5e-2


def _ray_samples(n, R=97, seed=0):
    rng = np.random.default_rng(seed)
    o = (rng.normal(size=(R, 3)) * 5 + [0, 0, 1500]).astype(np.float32)
    d = (rng.normal(size=(R, 3)) * 0.05 + [0, 0, -1]).astype(np.float32)
    ri = np.sort(rng.integers(0, R, n)).astype(np.int32)
    t0 = (1400 + rng.random(n) * 199).astype(np.float32); t1 = (t0 + 2.0 / 3.0).astype(np.float32)
    kw = dict(rays_o=torch.from_numpy(o).cuda(), rays_d=torch.from_numpy(d).cuda(), ray_idx=torch.from_numpy(ri).cuda(),
              t_starts=torch.from_numpy(t0).cuda(), t_ends=torch.from_numpy(t1).cuda())
    return o, d, ri, t0, t1, kw


def test_mlp256_ray_samples_alpha_ndev_and_index_list(A):
    p = ocppn.init_params(8, 256, "fourier", 5, 5.0, seed=2)
    model = _model(A, p, 8, "fourier", "bf16")
    n = 20000
    o, d, ri, t0, t1, kw = _ray_samples(n)
    pos = pipeline.midpoints(torch.from_numpy(o), torch.from_numpy(d), ri, torch.from_numpy(t0)[:, None], torch.from_numpy(t1)[:, None])
    with torch.no_grad():
        logit = ocppn.cppn_forward(p, pos, "fourier", 5).reshape(-1)
        alpha = (1 - torch.exp(-torch.sigmoid(logit) * torch.from_numpy(t1 - t0))).numpy()
    got = model.query(A.ops.OUT_ALPHA, **kw)
    assert np.abs(got.cpu().numpy() - alpha).max() <= 1.5e-2
    n_dev = torch.tensor([7001], dtype=torch.int32, device="cuda")
    got_d = model.query(A.ops.OUT_ALPHA, n_dev=n_dev, **kw)
    assert got_d[:7001].equal(got[:7001])
    ids = torch.arange(3, n, 7, dtype=torch.int32, device="cuda")
    model._ensure_flat()
    packed = A.ops.mlp_pack(model._desc, model._flat)
    out = torch.full((n,), -1.0, device="cuda")
    A.ops.mlp_forward(model._desc, model._flat, packed, A.ops.OUT_ALPHA, A.ops.PREC_BF16, out=out, sample_idx=ids, **kw)
    assert out[ids.long()].equal(got[ids.long()]) and float(out[0]) == -1.0


def _decode_saved_256(saved_u8, n, L, basis):
    """bf16 swizzled tile images (csrc/mlp_tc256.cu) -> the fp32 path's saved layout [X0 | a_1 | ... | a_{L+1}]"""
    n_tiles = (n + 127) // 128
    raw = saved_u8.cpu().numpy()
    r = np.arange(128)[:, None]; c = np.arange(64)[None, :]
    byte = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2

    def block(off):                                   # one [128, 64] block at byte offset off
        u16 = raw[off:off + 16384].view(np.uint16)[(byte // 2)]
        return (u16.astype(np.uint32) << 16).view(np.float32)

    a0 = np.concatenate([block(t * 16384) for t in range(n_tiles)])
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
        act = np.concatenate([np.concatenate([block(base + (l * n_tiles + t) * 65536 + b * 16384) for b in range(4)], axis=1)
                              for t in range(n_tiles)])[:n]
        buf = np.zeros(al(n * 256 * 4), np.uint8)
        buf[:n * 256 * 4] = np.ascontiguousarray(act).reshape(-1).view(np.uint8)
        parts.append(buf)
    return torch.from_numpy(np.concatenate(parts)).cuda()


@pytest.mark.parametrize("pos_enc,L,n", [("fourier", 8, 128 * 148 + 300), ("none", 2, 3000), ("fourier", 1, 100)])
def test_mlp256_backward(A, pos_enc, L, n):
    """tcgen05 forward(train) + dgrad + wgrad at width 256 against the fp32 check path on the same parameters and samples"""
    p = ocppn.init_params(L, 256, pos_enc, 5, 0.2, seed=4)
    model = _model(A, p, L, pos_enc, "bf16")
    model._ensure_flat()
    desc, flat = model._desc, model._flat
    packed = A.ops.mlp_pack(desc, flat)
    _, _, _, _, _, kw = _ray_samples(n, seed=n)
    g = torch.from_numpy(np.random.default_rng(n).normal(size=n).astype(np.float32)).cuda()
    y32, s32 = A.ops.mlp_forward(desc, flat, None, A.ops.OUT_LOGIT, A.ops.PREC_FP32, saved=True, **kw)
    g32 = A.ops.mlp_backward(desc, flat, None, s32, g, A.ops.PREC_FP32, **kw)
    y16, s16 = A.ops.mlp_forward(desc, flat, packed, A.ops.OUT_LOGIT, A.ops.PREC_BF16, saved=True, **kw)
    y16i = A.ops.mlp_forward(desc, flat, packed, A.ops.OUT_LOGIT, A.ops.PREC_BF16, **kw)
    assert torch.equal(y16, y16i)                                      # training forward == inference forward, bit for bit
    assert float((y16 - y32).abs().max()) <= 5e-2 * max(1.0, float(y32.abs().max()))
    g16 = A.ops.mlp_backward(desc, flat, packed, s16, g, A.ops.PREC_BF16, **kw)
    names = [k for k, _ in model.named_parameters() if not k.startswith("img")]
    # (a) end to end against the fp32 path: bounds gross errors only (ReLU sign flips between bf16 and fp32 forward, see the
    # width-128 test)
    rel_e2e = {nm: float((g16[o_:o_ + c_] - g32[o_:o_ + c_]).norm() / g32[o_:o_ + c_].norm().clamp_min(1e-12))
               for (o_, c_, _), nm in zip(model._param_slices, names)}
    assert max(rel_e2e.values()) <= 0.3, rel_e2e
    # (b) the backward kernels themselves: fp32 backward on the SAME (bf16) activations the tensor-core forward saved
    s32b = _decode_saved_256(s16, n, L, 5 if pos_enc == "fourier" else 0)
    g32b = A.ops.mlp_backward(desc, flat, None, s32b, g, A.ops.PREC_FP32, **kw)
    rels = {nm: round(float((g16[o_:o_ + c_] - g32b[o_:o_ + c_]).norm() / g32b[o_:o_ + c_].norm().clamp_min(1e-12)), 4)
            for (o_, c_, _), nm in zip(model._param_slices, names)}
    print(rels)
    assert max(rels.values()) <= 2.5e-2, rels


def _small_scene(A, res=32, W=24, seed=0):
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


def test_render_rays_8x256_vs_oracle(A):
    og, gg, roi, o, d = _small_scene(A)
    p = ocppn.init_params(8, 256, "fourier", 5, 0.05, seed=5)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] - 3.0
    f = functools.partial(ocppn.cppn_forward, p, pos_enc="fourier", basis=5)
    with torch.no_grad():
        pix_ref, (ri, ts, te) = pipeline.render_rays(f, og, roi, o, d, 300, 1400.0, 1600.0, 1e-2, 1e-4)
    pix_ref = pix_ref.numpy()
    model = _model(A, p, 8, "fourier", "bf16")
    with torch.no_grad():
        pix, (gi, g0, g1) = A.render_rays(model, gg, torch.tensor(roi).cuda(), torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(),
                                          300, 1400.0, 1600.0, 1e-2, 1e-4)
    pix = pix.cpu().numpy()
    assert np.max(np.abs(pix - pix_ref)) <= 1e-2 * pix_ref.max()            # north_star: bf16 projection <= 1e-2
    assert np.linalg.norm(pix - pix_ref) / np.linalg.norm(pix_ref) <= 1e-2


def test_training_step_8x256_vs_oracle_autograd(A):
    """one reference iteration with the 8 x 256 network on the tensor-core path: loss 5e-3, whole gradient <= 4e-2 relative L2
    against the oracle's torch autograd; the optimiser moves every parameter"""
    from nerf_for_angiography_b200.data import RayPool
    from nerf_for_angiography_b200.train import Trainer
    og, gg, roi, o, d = _small_scene(A, seed=3)
    L = 8
    p = ocppn.init_params(L, 256, "fourier", 5, 0.05, seed=7)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] - 3.0
    target = np.random.default_rng(1).random(len(o)).astype(np.float32)
    params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f = functools.partial(ocppn.cppn_forward, params, pos_enc="fourier", basis=5)
    pix_ref, (ri, ts, te) = pipeline.render_rays(f, og, roi, o, d, 300, 1400.0, 1600.0, 1e-2, 1e-4)
    loss_ref = torch.nn.functional.mse_loss(pix_ref, torch.from_numpy(target))
    loss_ref.backward()
    model = _model(A, p, L, "fourier", "bf16")
    pool = RayPool(torch.eye(4, dtype=torch.float64).cuda()[None], torch.zeros(1, 2, 2).cuda(), 1.0)
    tr = Trainer(model, pool, 1400.0, 1600.0, n_rays=len(o), vessel_grid=False)
    tr.acc_grid = gg
    tr.n_iter = 1
    before = tr.flat.clone()
    out = tr.step(rays=(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(target).cuda()))
    assert abs(out["n_samples"] - len(ri)) <= max(4, int(2e-3 * len(ri)))
    assert np.isclose(float(out["loss"]), float(loss_ref), rtol=5e-3)
    order = ["fourier_coefficients"] + [f"early_pts_layers.{2 * i}.{w}" for i in range(L + 1) for w in ("weight", "bias")] + \
        ["output_linear.0.weight", "output_linear.0.bias"]
    gref = np.concatenate([params[k].grad.numpy().reshape(-1) for k in order])
    got = tr.grad[:-1].cpu().numpy()
    rel = float(np.linalg.norm(got - gref) / np.linalg.norm(gref))
    print("8x256 bf16 gradient relative L2 vs oracle autograd:", round(rel, 4))
    assert rel <= 4e-2, rel
    assert float((tr.flat - before).abs().max()) > 0


def test_chunked_backward_equals_one_shot(A):
    """one-sync path with the saved tile images over budget: forward(train) / backward run chunk by chunk with the gradient
    accumulated -- same loss, same gradient (fp32 summation order of the weight gradients aside) as the one-shot step"""
    from nerf_for_angiography_b200.data import RayPool
    from nerf_for_angiography_b200.train import Trainer
    og, gg, roi, o, d = _small_scene(A, seed=4)
    p = ocppn.init_params(2, 256, "fourier", 5, 0.05, seed=9)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] - 3.0
    target = torch.from_numpy(np.random.default_rng(3).random(len(o)).astype(np.float32)).cuda()
    rays = (torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), target)
    res = []
    for budget in (None, 1):
        model = _model(A, p, 2, "fourier", "bf16")
        pool = RayPool(torch.eye(4, dtype=torch.float64).cuda()[None], torch.zeros(1, 2, 2).cuda(), 1.0)
        tr = Trainer(model, pool, 1400.0, 1600.0, n_rays=len(o), vessel_grid=False, sync_free=False)
        tr.acc_grid = gg
        tr.n_iter = 1
        if budget is not None:
            tr.train_memory_bytes = budget                      # forces the smallest chunk (128 x 148 samples)
        out = tr.step(rays=rays)
        if budget is not None:
            assert tr._backward_chunk(out["n_samples"]) == 128 * 148 < out["n_samples"]
        res.append((float(out["loss"]), tr.grad[:-1].clone(), out["n_samples"]))
    assert res[0][2] == res[1][2] and np.isclose(res[0][0], res[1][0], rtol=1e-6)
    scale = float(res[0][1].abs().max())
    assert float((res[0][1] - res[1][1]).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("precision,tol_y,tol_g", [("fp32", 1e-5, 1e-4), ("bf16", 3e-2, 2e-1)])
def test_cppn_width256_matches_reference_golden(A, golden_dir, precision, tol_y, tol_g):
    """golden vectors produced by the REFERENCE's own CPPN class at hidden width 256 (tests/golden/make_golden.py,
    cppn_fourier_2x256.npz: state dict, 192 points in +-100, output, parameter gradients): the CPPN module drop-in on the fp32
    check path (1e-5 / 1e-4) and on the width-256 tcgen05 path (bf16: 3e-2 of the output scale; gradients over only 192 points are
    dominated by the few ReLU units whose sign differs between bf16 and fp32, so they only bound gross errors here: 2e-1 relative L2,
    like the width-128 BARF golden; the tight gradient checks are test_mlp256_backward and the training-step test)"""
    import os
    g = np.load(os.path.join(golden_dir, "cppn_fourier_2x256.npz"))
    model = A.CPPN(_mdef(2, "fourier", precision))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    assert set(model.state_dict().keys()) == set(sd.keys())          # reference checkpoint keys, verbatim
    model.load_state_dict(sd)
    model = model.to("cuda")
    assert model.precision == precision
    x = torch.from_numpy(g["x"]).cuda()
    y = model(x)
    ref = g["y"]
    assert np.max(np.abs(y.detach().cpu().numpy() - ref)) <= tol_y * max(1.0, np.abs(ref).max())
    (y * torch.from_numpy(g["gout"]).cuda()).sum().backward()
    for name, p in model.named_parameters():
        if "grad:" + name in g.files:
            gr = g["grad:" + name]
            assert p.grad is not None, name
            got = p.grad.cpu().numpy()
            if precision == "fp32":
                assert np.max(np.abs(got - gr)) <= tol_g * max(np.abs(gr).max(), 1e-6), name
            else:
                assert np.linalg.norm(got - gr) <= tol_g * max(np.linalg.norm(gr), 1e-6), (name, np.linalg.norm(got - gr) / np.linalg.norm(gr))
