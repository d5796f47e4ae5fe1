"""GPU parity tests: every kernel group of libangio_b200.so, called through the C ABI, against the CPU oracle and
the golden vectors generated from the reference's own code.  Integer / index outputs are compared bit-exactly,
floating point within the tolerance written next to each assert."""
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
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return a


def _dev(x, dtype=None):
    t = torch.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


# ------------------------------------------------------------------------------------------------ ray generation
def test_raygen_bit_exact_vs_reference_golden(A, golden_dir):
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    for ci in range(int(g["n_cases"])):
        th, ph, la, W, H, f, tx, ty, tz = g[f"c{ci}_args"]
        o, d, M, ii, jj = A.get_ray_values(th, ph, la, np.array([0.0, 0.0, 1500.0]), int(W), int(H), f, "cuda",
                                           np.array([tx, ty, tz]))
        assert np.array_equal(M, g[f"c{ci}_M"])
        assert o.shape == (int(H), int(W), 3) and o.dtype == torch.float32
        # the reference casts its float64 rays with .float() (run_nerf_acc.py:88-89): compare bit for bit
        assert np.array_equal(o.cpu().numpy(), g[f"c{ci}_o"].astype(np.float32))
        assert np.array_equal(d.cpu().numpy(), g[f"c{ci}_d"].astype(np.float32))


def test_raygen_gather_mode(A):
    mats = np.stack([ogeo.source_matrix([0, 0, 1500.0], th, 0.0) for th in (0.0, 30.0, 77.0)])
    W, H, f = 20, 12, 150.0
    rng = np.random.default_rng(0)
    v = rng.integers(0, 3, 500).astype(np.int32)
    x = rng.integers(0, W, 500).astype(np.int32)
    y = rng.integers(0, H, 500).astype(np.int32)
    pix = rng.random((3, H, W), dtype=np.float32)
    o, d, p = A.ops.raygen(_dev(mats), W, H, f, view_ids=_dev(v), px=_dev(x), py=_dev(y), pixels=_dev(pix))
    for k, th in enumerate((0.0, 30.0, 77.0)):
        ro, rd, _ = ogeo.get_ray_values(th, 0.0, 0.0, [0, 0, 1500.0], W, H, f)
        sel = v == k
        assert np.array_equal(d.cpu().numpy()[sel], rd[y[sel], x[sel]].astype(np.float32))
        assert np.array_equal(o.cpu().numpy()[sel], ro[y[sel], x[sel]].astype(np.float32))
    assert np.array_equal(p.cpu().numpy(), pix[v, y, x])
    # flat-id mode (the sampler's output feeds it directly): same rays bit for bit
    ids = torch.from_numpy((v.astype(np.int64) * H + y) * W + x).cuda()
    o2, d2, p2 = A.ops.raygen_flat(_dev(mats), ids, W, H, f, pixels=_dev(pix))
    assert o2.equal(o) and d2.equal(d) and p2.equal(p)


# ------------------------------------------------------------------------------------------------ marching
def _scene(n_rays=700, W=32, seed=0, views=((0.0, 0.0), (45.0, 0.0), (135.0, 135.0))):
    rng = np.random.default_rng(seed)
    os_, ds_ = [], []
    for th, ph in views:
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], W, W, 7.5 * W)
        os_.append(o.reshape(-1, 3)); ds_.append(d.reshape(-1, 3))
    o = np.concatenate(os_).astype(np.float32); d = np.concatenate(ds_).astype(np.float32)
    sel = rng.choice(len(o), n_rays, replace=False)
    return o[sel], d[sel]


def _march_both(A, o, d, binary, res, near=1400.0, far=1600.0, n_steps=300, aabb=None):
    aabb = np.array([-100, -100, -100, 100, 100, 100], np.float32) if aabb is None else aabb
    step = np.float32((far - near) / n_steps)
    tmin, tmax = nerfacc_ref.ray_aabb_intersect(o, d, aabb, near, far)
    ri, ts, te, off = nerfacc_ref.march(o, d, tmin, tmax, aabb, res, binary, step)
    gi, g0, g1, goff = A.ops.march(_dev(o), _dev(d), aabb, aabb, res, _dev(binary), near, far, float(step))
    # the write pass replays the runs recorded by the count pass (default) or marches the grid a second time: same samples
    ni, n0, n1, noff = A.ops.march(_dev(o), _dev(d), aabb, aabb, res, _dev(binary), near, far, float(step), use_runs=False)
    assert ni.equal(gi) and n0.equal(g0) and n1.equal(g1) and noff.equal(goff)
    # capacity mode (no host sync): same samples in the first offsets[R] slots, count left on the device
    cap = A.ops.march_capacity(len(o), near, far, float(step))
    tot = torch.zeros(1, dtype=torch.int32, device="cuda")
    ci, c0, c1, coff = A.ops.march(_dev(o), _dev(d), aabb, aabb, res, _dev(binary), near, far, float(step), capacity=cap, total_out=tot)
    n_c = int(tot.item())
    assert n_c == gi.numel() <= cap and coff.equal(goff)
    assert ci[:n_c].equal(gi) and c0[:n_c].equal(g0) and c1[:n_c].equal(g1)
    return (ri, ts, te, off), (gi.cpu().numpy(), g0.cpu().numpy(), g1.cpu().numpy(), goff.cpu().numpy())


@pytest.mark.parametrize("fill", ["random", "full", "empty", "single", "blobs"])
def test_march_bit_exact(A, fill):
    res = 32
    rng = np.random.default_rng(3)
    if fill == "random":
        binary = rng.random((res,) * 3) < 0.3
    elif fill == "full":
        binary = np.ones((res,) * 3, bool)
    elif fill == "empty":
        binary = np.zeros((res,) * 3, bool)
    elif fill == "single":
        binary = np.zeros((res,) * 3, bool); binary[16, 15, 17] = True
    else:
        r = np.arange(res)
        X, Y, Z = np.meshgrid(r, r, r, indexing="ij")
        binary = ((X - 10) ** 2 + (Y - 20) ** 2 + (Z - 14) ** 2 < 40) | ((X - 22) ** 2 + (Y - 9) ** 2 + (Z - 20) ** 2 < 25)
    o, d = _scene()
    (ri, ts, te, off), (gi, g0, g1, goff) = _march_both(A, o, d, binary, res)
    assert np.array_equal(goff.astype(np.int64), off)                 # segment offsets: bit-exact
    assert np.array_equal(gi.astype(np.int64), ri)                    # ray indices: bit-exact
    assert np.array_equal(g0, ts) and np.array_equal(g1, te)          # fp32 interval ends: bit-exact
    if fill == "empty":
        assert len(gi) == 0
    if fill == "full":
        assert len(gi) > 100 * len(o) * 0.5


@pytest.mark.parametrize("dist,n_steps", [(1000.0, 300), (1000.0, 77), (60.0, 300), (2100.0, 500)])
def test_march_bit_exact_across_binades(A, dist, n_steps, monkeypatch):
    """The warp-per-ray count / write kernels use the closed form of the t-chain only for rays whose whole chain lies in one fp32
    binade; at these source distances [t_min, t_max] straddles 1024 / 64 / 2048, so some rays take the closed form and the others
    the serial walk -- all of them bit-equal to the oracle, and to the all-serial kernels (ANGIO_MARCH_SERIAL_*)."""
    res = 32
    rng = np.random.default_rng(11)
    binary = rng.random((res,) * 3) < 0.35
    W = 32
    half = 100.0 if dist > 200 else 30.0
    os_, ds_ = [], []
    for th, ph in ((0.0, 0.0), (45.0, 10.0), (135.0, 135.0)):
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, dist], W, W, (W / 2) * dist / half)
        os_.append(o.reshape(-1, 3)); ds_.append(d.reshape(-1, 3))
    o = np.concatenate(os_).astype(np.float32); d = np.concatenate(ds_).astype(np.float32)
    aabb = np.array([-half] * 3 + [half] * 3, np.float32)
    near, far = dist - half, dist + half
    (ri, ts, te, off), (gi, g0, g1, goff) = _march_both(A, o, d, binary, res, near=near, far=far, n_steps=n_steps, aabb=aabb)
    assert np.array_equal(goff.astype(np.int64), off) and np.array_equal(gi.astype(np.int64), ri)
    assert np.array_equal(g0, ts) and np.array_equal(g1, te)
    assert len(gi) > 0
    monkeypatch.setenv("ANGIO_MARCH_SERIAL_COUNT", "1")
    monkeypatch.setenv("ANGIO_MARCH_SERIAL_WRITE", "1")
    step = np.float32((far - near) / n_steps)
    si, s0, s1, soff = A.ops.march(_dev(o), _dev(d), aabb, aabb, res, _dev(binary), near, far, float(step))
    assert np.array_equal(si.cpu().numpy(), gi) and np.array_equal(s0.cpu().numpy(), g0) and np.array_equal(s1.cpu().numpy(), g1)


def test_march_degenerate_rays(A):
    """axis-parallel directions (zero components -> inf / NaN in the slab test), rays missing the box, zero rays"""
    res = 16
    binary = np.ones((res,) * 3, bool)
    o = np.array([[0, 0, 1500], [0, 0, 1500], [300, 0, 1500], [0, 0, 1500], [5, -3, 1500]], np.float32)
    d = np.array([[0, 0, -1], [0.01, 0, -1], [0, 0, -1], [0, 1, 0], [-0.0, 0.0, -1]], np.float32)
    (ri, ts, te, off), (gi, g0, g1, goff) = _march_both(A, o, d, binary, res)
    assert np.array_equal(goff.astype(np.int64), off) and np.array_equal(g0, ts) and np.array_equal(g1, te)
    assert off[3] - off[2] == 0                                        # the ray that misses the box has no samples
    e = A.ops.march(torch.zeros((0, 3), device="cuda"), torch.zeros((0, 3), device="cuda"), np.array([-1, -1, -1, 1, 1, 1.0]),
                    np.array([-1, -1, -1, 1, 1, 1.0]), res, _dev(binary), 0.0, 1.0, 0.1)
    assert e[0].numel() == 0 and e[3].tolist() == [0]


def test_march_known_answer_full_grid_central_ray(A):
    """central ray of the theta=0 view: enters the +-100 box at t=1400, leaves at 1600 -> exactly 300 samples of
    width 2/3 starting at near (analytic known answer)"""
    res = 128
    binary = np.ones((res,) * 3, bool)
    o = np.array([[0, 0, 1500.0]], np.float32); d = np.array([[0, 0, -1.0]], np.float32)
    _, (gi, g0, g1, goff) = _march_both(A, o, d, binary, res)
    assert abs(len(gi) - 300) <= 1
    assert g0[0] == np.float32(1400.0)
    assert np.allclose(g1 - g0, 2.0 / 3.0, atol=1e-3)


def test_grid_query(A):
    res = 16
    rng = np.random.default_rng(5)
    binary = rng.random((res,) * 3) < 0.5
    pts = (rng.random((4000, 3), dtype=np.float32) * 2 - 1) * 120
    g = nerfacc_ref.OccupancyGrid([-100, -100, -100, 100, 100, 100], res); g.binary = binary
    got = A.ops.grid_query(_dev(pts), g.roi_aabb, res, _dev(binary))
    assert np.array_equal(got.cpu().numpy(), g.query_occ(pts))


@pytest.mark.parametrize("res,lo,hi", [(128, -100.0, 100.0), (16, -100.0, 100.0), (96, -37.5, 81.25)])
def test_cell_index_is_bit_exact_at_cell_boundaries(A, res, lo, hi):
    """Cell lookup at every cell boundary of every axis, +-0..40 ulps, against the reference formula
    trunc(fl(fl(x - lo) / ext) * res) evaluated in numpy float32 -- a checkerboard grid makes any off-by-one cell visible in the
    occupancy answer.  (A multiply-by-reciprocal shortcut with a guard band passed this test too but was slower than the IEEE
    division it replaced, 177 vs 153 us for the count pass, and was dropped.)"""
    f32 = np.float32
    ext = f32(f32(hi) - f32(lo))
    xs = []
    for k in range(res + 1):
        b = f32(f32(lo) + f32(k) * ext / f32(res))
        v = b
        for _ in range(40):
            v = np.nextafter(v, f32(-np.inf), dtype=f32)
        for _ in range(81):
            xs.append(v); v = np.nextafter(v, f32(np.inf), dtype=f32)
    xs = np.array(xs, f32)
    rng = np.random.default_rng(res)
    mid = f32((lo + hi) / 2)
    pts = np.stack([np.stack([xs, np.full_like(xs, mid), np.full_like(xs, mid)], 1),
                    np.stack([np.full_like(xs, mid), xs, np.full_like(xs, mid)], 1),
                    np.stack([np.full_like(xs, mid), np.full_like(xs, mid), xs], 1)]).reshape(-1, 3)
    pts = np.concatenate([pts, (rng.random((20000, 3), dtype=f32) * f32(ext) + f32(lo)).astype(f32)])
    I, J, K = np.meshgrid(np.arange(res), np.arange(res), np.arange(res), indexing="ij")
    binary = ((I + J + K) % 2 == 0)

    def ref_index(x):
        a = (x - f32(lo)).astype(f32)
        return np.clip(((a / ext).astype(f32) * f32(res)).astype(f32).astype(np.int32), 0, res - 1)
    inside = np.all((pts >= f32(lo)) & (pts <= f32(hi)), axis=1)
    want = np.where(inside, binary[ref_index(pts[:, 0]), ref_index(pts[:, 1]), ref_index(pts[:, 2])], False).astype(np.float32)
    roi = np.array([lo] * 3 + [hi] * 3, f32)
    got = A.ops.grid_query(_dev(pts), roi, res, _dev(binary)).cpu().numpy()
    assert np.array_equal(got, want), int((got != want).sum())


# ------------------------------------------------------------------------------------------------ visibility + compaction
@pytest.mark.parametrize("thre", [0.0, 1e-2])
def test_visibility_compaction_bit_exact(A, thre):
    rng = np.random.default_rng(7)
    R = 900
    counts = rng.integers(0, 90, R); counts[::17] = 0
    off = np.zeros(R + 1, np.int64); np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    alphas = (rng.random(n, dtype=np.float32) ** 3 * 0.4).astype(np.float32)
    t0 = rng.random(n, dtype=np.float32) * 100 + 1400; t1 = t0 + 0.66
    keep = nerfacc_ref.visibility(off, alphas, 1e-2, thre)
    ridx = np.repeat(np.arange(R), counts)
    gi, g0, g1, goff, gkeep = A.ops.visibility_compact(_dev(alphas), _dev(off, torch.int32), _dev(t0), _dev(t1), 1e-2, thre)
    assert np.array_equal(gkeep.cpu().numpy().astype(bool), keep)
    assert np.array_equal(gi.cpu().numpy(), ridx[keep])
    assert np.array_equal(g0.cpu().numpy(), t0[keep]) and np.array_equal(g1.cpu().numpy(), t1[keep])
    new_counts = np.bincount(ridx[keep], minlength=R)
    assert np.array_equal(np.diff(goff.cpu().numpy()), new_counts)


# ------------------------------------------------------------------------------------------------ composite
def test_composite_vs_reference_golden(A, golden_dir):
    g = np.load(os.path.join(golden_dir, "composite.npz"))
    n_rays = int(g["n_rays"])
    pred = _dev(g["pred"]).requires_grad_(True)
    ri = _dev(g["ray_indices"])
    pix, ent = A.acc_render_volume_density(pred, ri, _dev(g["t_starts"]), _dev(g["t_ends"]), n_rays, 300)
    assert ent is None and pix.shape == (n_rays,) and pix.dtype == torch.float32
    assert np.allclose(pix.detach().cpu().numpy(), g["pix"], rtol=2e-6, atol=1e-7)      # fp32 tolerance 2e-6
    assert float(pix[5]) == 1.0 and float(pix[36]) == 1.0                               # empty rays render exactly 1
    (pix * _dev(g["gpix"])).sum().backward()
    assert np.allclose(pred.grad.cpu().numpy(), g["gpred"], rtol=1e-5, atol=1e-7)
    zero_idx = torch.where(torch.sigmoid(pred.detach()) < 0.4)
    pz, _ = A.acc_render_volume_density(pred.detach(), ri, _dev(g["t_starts"]), _dev(g["t_ends"]), n_rays, 300, zero_idx)
    assert np.allclose(pz.cpu().numpy(), g["pix_zero"], rtol=2e-6, atol=1e-7)


def test_composite_mse_fused_matches_autograd_oracle(A):
    rng = np.random.default_rng(11)
    R = 257
    counts = rng.integers(0, 70, R)
    off = np.zeros(R + 1, np.int64); np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    ri = np.repeat(np.arange(R), counts)
    logits = rng.normal(size=n).astype(np.float32)
    t0 = (rng.random(n) * 100 + 1400).astype(np.float32); t1 = (t0 + 0.667).astype(np.float32)
    target = rng.random(R).astype(np.float32)
    p = torch.from_numpy(logits)[:, None].clone().requires_grad_(True)
    pix = pipeline.acc_render_volume_density(p, ri, torch.from_numpy(t0)[:, None], torch.from_numpy(t1)[:, None], R)
    loss = torch.nn.functional.mse_loss(pix, torch.from_numpy(target))
    loss.backward()
    gp, gg, gl = A.ops.composite_mse_fused(_dev(logits), _dev(t0), _dev(t1), _dev(off, torch.int32), _dev(target))
    assert np.allclose(gp.cpu().numpy(), pix.detach().numpy(), rtol=2e-6, atol=1e-7)
    assert np.isclose(float(gl) / R, float(loss), rtol=1e-5)
    assert np.allclose(gg.cpu().numpy(), p.grad.numpy().reshape(-1), rtol=1e-4, atol=1e-9)


# ------------------------------------------------------------------------------------------------ MLP (fp32 check mode)
def _flat_from_sd(sd, enc):
    parts = []
    if enc:
        parts.append(sd["fourier_coefficients"].reshape(-1))
    i = 0
    while f"early_pts_layers.{2 * i}.weight" in sd:
        parts += [sd[f"early_pts_layers.{2 * i}.weight"].reshape(-1), sd[f"early_pts_layers.{2 * i}.bias"].reshape(-1)]
        i += 1
    parts += [sd["output_linear.0.weight"].reshape(-1), sd["output_linear.0.bias"].reshape(-1)]
    return np.concatenate(parts).astype(np.float32), i - 1


@pytest.mark.parametrize("tag,enc,width", [("none_2x64", 0, 64), ("fourier_4x128", 1, 128), ("fourier_2x64", 1, 64)])
def test_mlp_fp32_vs_reference_golden(A, golden_dir, tag, enc, width):
    g = np.load(os.path.join(golden_dir, f"cppn_{tag}.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd:")}
    flat, n_hidden = _flat_from_sd(sd, enc)
    desc = A.ops.mlp_desc(enc, 5 if enc else 0, width, n_hidden)
    assert A.ops.mlp_param_count(desc) == flat.size
    x = _dev(g["x"])
    y, saved = A.ops.mlp_forward(desc, _dev(flat), None, A.ops.OUT_LOGIT, A.ops.PREC_FP32, saved=True, points=x)
    ref = g["y"].reshape(-1)
    # fp32 check mode: <= 1e-5 relative to the output scale (north_star tolerance for the fp32 mode)
    assert np.max(np.abs(y.cpu().numpy() - ref)) <= 1e-5 * max(1.0, np.max(np.abs(ref)))
    grad = A.ops.mlp_backward(desc, _dev(flat), None, saved, _dev(g["gout"].reshape(-1)), A.ops.PREC_FP32, points=x)
    gsd = {k[5:]: g[k] for k in g.files if k.startswith("grad:")}
    gref, _ = _flat_from_sd(gsd, enc)
    got = grad.cpu().numpy()
    denom = np.max(np.abs(gref))
    assert np.max(np.abs(got - gref)) <= 2e-5 * denom, np.max(np.abs(got - gref)) / denom


def test_mlp_fp32_output_modes_and_ray_samples(A):
    p = ocppn.init_params(2, 64, "fourier", 5, 5.0, seed=3)
    flat, n_hidden = _flat_from_sd({k: v.numpy() for k, v in p.items()}, 1)
    desc = A.ops.mlp_desc(1, 5, 64, n_hidden)
    rng = np.random.default_rng(0)
    R, n = 50, 3000
    o = (rng.normal(size=(R, 3)) * 5 + [0, 0, 1500]).astype(np.float32)
    d = rng.normal(size=(R, 3)).astype(np.float32) * 0.1 + np.array([0, 0, -1], np.float32)
    ri = np.sort(rng.integers(0, R, n)).astype(np.int32)
    t0 = (1400 + rng.random(n) * 199).astype(np.float32); t1 = (t0 + 2.0 / 3.0).astype(np.float32)
    pos = pipeline.midpoints(torch.from_numpy(o), torch.from_numpy(d), ri, torch.from_numpy(t0)[:, None], torch.from_numpy(t1)[:, None])
    logit = ocppn.cppn_forward(p, pos, "fourier", 5).reshape(-1)
    sig = torch.sigmoid(logit)
    alpha = 1 - torch.exp(-sig * torch.from_numpy(t1 - t0))
    kw = dict(rays_o=_dev(o), rays_d=_dev(d), ray_idx=_dev(ri), t_starts=_dev(t0), t_ends=_dev(t1))
    f = functools.partial(A.ops.mlp_forward, desc, _dev(flat), None)
    assert np.allclose(f(A.ops.OUT_LOGIT, A.ops.PREC_FP32, **kw).cpu().numpy(), logit.numpy(), rtol=1e-4, atol=2e-5)
    assert np.allclose(f(A.ops.OUT_SIGMA, A.ops.PREC_FP32, **kw).cpu().numpy(), sig.numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(f(A.ops.OUT_ALPHA, A.ops.PREC_FP32, **kw).cpu().numpy(), alpha.numpy(), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------ occupancy grid
def test_grid_update_vs_oracle(A):
    res = 16
    roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
    rng = np.random.default_rng(2)
    og = nerfacc_ref.OccupancyGrid(roi, res)
    gg = A.OccupancyGrid(torch.tensor(roi), res, A.ContractionType.AABB).cuda()
    field = lambda x: (1.0 / (1.0 + np.exp(-(np.asarray(x)[:, 0] / 40.0 - 1.0)))).astype(np.float32) * 1e-2  # noqa: E731
    # warm-up update over all cells, then a sparse duplicate-free update
    for step, cells in [(0, None), (16, rng.permutation(res ** 3)[:1000].astype(np.int64))]:
        n = res ** 3 if cells is None else len(cells)
        jit = rng.random((n, 3), dtype=np.float32)
        idx = np.arange(res ** 3) if cells is None else cells
        x_ref = og.cell_points(idx, jit)
        x_gpu = A.ops.grid_cell_points(None if cells is None else _dev(cells), _dev(jit), roi, res)
        assert np.array_equal(x_gpu.cpu().numpy(), x_ref)                       # cell -> point map: bit-exact
        og.update_from_occ(idx, field(x_ref), occ_thre=5e-3)
        gg._update(step, lambda x: _dev(field(x.cpu().numpy())), occ_thre=5e-3, cells=None if cells is None else _dev(cells),
                   jitter=_dev(jit))
        assert np.array_equal(gg.occs.cpu().numpy(), og.occs)                   # EMA-max: bit-exact
        assert np.array_equal(gg.binary.cpu().numpy(), og.binary)
        assert np.isclose(gg.occs_mean_host, float(og.occs.mean(dtype=np.float32)), rtol=1e-5)
    # duplicates: kernel takes the max over duplicates of the once-decayed value
    cells = np.array([5, 5, 9], np.int64); occ = np.array([0.1, 0.3, 0.2], np.float32)
    before = gg.occs.cpu().numpy().copy()
    A.ops.grid_ema_update(gg.occs, _dev(cells), _dev(occ), 0.95)
    after = gg.occs.cpu().numpy()
    assert after[5] == max(np.float32(before[5] * np.float32(0.95)), np.float32(0.3)) and after[9] == max(np.float32(before[9] * np.float32(0.95)), np.float32(0.2))


def test_adam_matches_torch(A):
    torch.manual_seed(0)
    p = torch.randn(1000); g1 = torch.randn(1000); g2 = torch.randn(1000)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4)
    gp, m, v = p.clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    for step, g in enumerate([g1, g2], start=1):
        ref.grad = g.clone(); opt.step()
        A.ops.adam_step(gp, g.cuda(), m, v, 1e-4, step)
        assert torch.allclose(gp.cpu(), ref.detach(), rtol=1e-6, atol=1e-8)


# ------------------------------------------------------------------------------------------------ ray sampling
def test_weighted_sampling_without_replacement(A):
    from nerf_for_angiography_b200.data import RayPool
    V, H, W = 3, 40, 50
    torch.manual_seed(0)
    w = torch.rand(V, H, W, device="cuda") ** 2 + 1e-3
    w[1] *= 5.0                                             # view 1 five times as likely
    pool = RayPool(torch.eye(4, dtype=torch.float64, device="cuda").repeat(V, 1, 1), torch.rand(V, H, W, device="cuda"), 100.0, w)
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 1500
    counts = torch.zeros(V * H * W, device="cuda")
    for _ in range(60):
        ids = pool.sample_ids(n, generator=g)
        assert ids.numel() == n and ids.unique().numel() == n      # without replacement
        assert int(ids.min()) >= 0 and int(ids.max()) < V * H * W
        counts[ids] += 1
    # inclusion frequencies match numpy's sequential weighted sampling without replacement -- what pandas'
    # DataFrame.sample(n, weights=...) does (nerf/nerf_helpers.py:139) -- here at a 25 % sampling fraction where heavy rays saturate
    wn = w.reshape(-1).double().cpu().numpy(); wn /= wn.sum()
    rng = np.random.default_rng(0)
    ref = np.zeros(V * H * W)
    for _ in range(60):
        ref[rng.choice(V * H * W, n, replace=False, p=wn)] += 1
    view_got = counts.view(V, -1).sum(1).cpu().numpy()
    view_ref = ref.reshape(V, -1).sum(1)
    assert np.allclose(view_got, view_ref, rtol=0.03), (view_got, view_ref)
    # all rays requested -> a permutation of the pool
    ids = pool.sample_ids(V * H * W, generator=g)
    assert ids.sort().values.equal(torch.arange(V * H * W, device="cuda"))
    assert pool.last_status.tolist()[1] == 0
    # uniform pool
    pool_u = RayPool(pool.cam2world, pool.pixels, 100.0, None)
    ids = pool_u.sample_ids(1000, generator=g)
    assert ids.unique().numel() == 1000


def test_sampler_edge_cases(A):
    """n = 1, a pool where almost every weight is zero (only positive-weight rays may be drawn), and n = every positive ray."""
    from nerf_for_angiography_b200.data import RayPool
    V, H, W = 2, 64, 64
    w = torch.zeros(V, H, W, device="cuda")
    pos = torch.randperm(V * H * W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))[:300]
    w.view(-1)[pos] = torch.rand(300, device="cuda") + 0.1
    pool = RayPool(torch.eye(4, dtype=torch.float64, device="cuda").repeat(V, 1, 1), torch.rand(V, H, W, device="cuda"), 100.0, w)
    g = torch.Generator(device="cuda").manual_seed(5)
    one = pool.sample_ids(1, generator=g)
    assert one.numel() == 1 and float(w.view(-1)[one]) > 0 and pool.last_status.tolist()[1] == 0
    some = pool.sample_ids(100, generator=g)
    assert some.unique().numel() == 100 and bool((w.view(-1)[some] > 0).all())
    allpos = pool.sample_ids(300, generator=g)
    assert allpos.sort().values.equal(pos.sort().values) and pool.last_status.tolist()[1] == 0
    # asking for more rays than have positive weight cannot be satisfied: the pool refuses (numpy / pandas raise here too) ...
    with pytest.raises(ValueError):
        pool.sample_ids(301, generator=g)
    # ... and the kernels themselves flag the failed draw and leave VALID ids behind (the gather that follows reads them
    # before any host code has looked at the status)
    wf = w.reshape(-1).contiguous()
    ids, status = A.ops.sample_without_replacement(301, V * H * W, wf, float(wf.sum()), float((wf.double() ** 2).sum()), 7, wf.device)
    assert status.tolist()[1] == 1 and int(ids.min()) == 0 and int(ids.max()) == 0


def test_sampler_with_the_reference_weight_images(A):
    """make_dataset(weight_strategy='distance'): the reference's mask -> EDT weight images (background pixels weigh 1e-10,
    helpers.py:226-247) through the sampler: the draw is complete, unique and lands on the weighted pixels."""
    from nerf_for_angiography_b200.data import make_dataset
    pool, _ = make_dataset(img_size=32, thetas=(0.0, 60.0, 120.0), kind="ct", volume_res=32, device="cuda", weight_strategy="distance")
    w = pool.weights.reshape(-1)
    assert float(w.min()) > 0 and abs(float(w.max()) - 1.0) < 1e-6 and int((w > 1e-6).sum()) > 600
    ids = pool.sample_ids(512, generator=torch.Generator(device="cuda").manual_seed(3))
    assert pool.last_status.tolist()[1] == 0 and ids.unique().numel() == 512
    assert float((w[ids] > 1e-6).float().mean()) > 0.99


def test_pool_can_exclude_the_test_view(A):
    """The reference trains on the test view as well (run_nerf_acc.py:114) -- the default; train_on_test_view=False keeps the
    last view out of the training draws (uniform and weighted)."""
    from nerf_for_angiography_b200.data import RayPool
    V, H, W = 3, 32, 48
    cam = torch.eye(4, dtype=torch.float64, device="cuda").repeat(V, 1, 1)
    w = torch.rand(V, H, W, device="cuda") + 0.1
    w[-1] = 100.0                                                 # the test view would dominate a weighted draw
    for weights in (None, w):
        ref = RayPool(cam, torch.rand(V, H, W, device="cuda"), 100.0, weights)
        ids = ref.sample_ids(2000, generator=torch.Generator(device="cuda").manual_seed(1))
        assert int(ids.max()) >= (V - 1) * H * W                  # reference behaviour: test-view rays are drawn
        pool = RayPool(cam, torch.rand(V, H, W, device="cuda"), 100.0, weights, train_on_test_view=False)
        ids = pool.sample_ids(2000, generator=torch.Generator(device="cuda").manual_seed(1))
        assert pool.last_status.tolist()[1] == 0 and ids.unique().numel() == 2000 and int(ids.max()) < (V - 1) * H * W
        assert pool.rays_of_view(V - 1)[2].numel() == H * W       # the test view itself stays available for evaluation
    with pytest.raises(ValueError):
        pool.sample_ids((V - 1) * H * W + 1)


def test_sampler_is_reproducible_and_shuffled(A):
    """Same seed -> the same ids in the same order (the candidate pass appends with atomics, the select/shuffle pass must
    erase that order); the order is a uniform shuffle (no correlation between position and ray id or weight)."""
    from nerf_for_angiography_b200.data import RayPool
    V, H, W = 8, 256, 256
    torch.manual_seed(1)
    w = torch.rand(V, H, W, device="cuda") + 0.05
    pool = RayPool(torch.eye(4, dtype=torch.float64, device="cuda").repeat(V, 1, 1), torch.rand(V, H, W, device="cuda"), 100.0, w)
    n = 8192
    runs = []
    for _ in range(3):
        g = torch.Generator(device="cuda").manual_seed(11)
        pool._seed_streams = {}
        runs.append(pool.sample_ids(n, generator=g))
    assert runs[0].equal(runs[1]) and runs[0].equal(runs[2])
    ids = runs[0]
    assert ids.unique().numel() == n and pool.last_status.tolist()[1] == 0
    pos = torch.arange(n, device="cuda", dtype=torch.float64)
    for other in (ids.double(), w.reshape(-1)[ids].double()):
        c = torch.corrcoef(torch.stack([pos, other]))[0, 1]
        assert abs(float(c)) < 0.05, float(c)                  # |corr| ~ 1/sqrt(n) = 0.011 for a uniform shuffle
    # exactness of the selection: the sample is the n smallest keys <=> inclusion probability grows with the weight
    heavy = w.reshape(-1)[ids].mean()
    assert float(heavy) > float(w.mean()) * 1.15
