"""CPU-only: analytic known-answer tests and properties of the oracle's nerfacc restatement."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import geometry as ogeo, nerfacc_ref, pipeline

ROI = np.array([-100, -100, -100, 100, 100, 100], np.float32)


def _rays(theta=0.0, phi=0.0, W=16):
    o, d, _ = ogeo.get_ray_values(theta, phi, 0.0, [0, 0, 1500.0], W, W, 7.5 * W)
    return o.reshape(-1, 3).astype(np.float32), d.reshape(-1, 3).astype(np.float32)


def _grid(res, fill):
    g = nerfacc_ref.OccupancyGrid(ROI, res)
    g.binary[:] = fill
    g.occs[:] = 0.5 if fill else 0.0
    return g


def test_central_ray_passes_through_the_origin():
    for th, ph in [(0, 0), (45, 0), (135, 135), (354, 0)]:
        o, d, M = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], 8, 8, 60.0)
        c = o[4, 4] + 1500.0 * d[4, 4]          # pixel (W/2, H/2) is the principal ray
        assert np.allclose(c, 0.0, atol=1e-9)
        assert np.isclose(np.linalg.norm(o[0, 0]), 1500.0)


def test_empty_grid_gives_no_samples_and_white_pixels():
    o, d = _rays()
    g = _grid(16, False)
    pix, (ri, ts, te) = pipeline.render_rays(lambda x: torch.zeros(len(x), 1), g, ROI, o, d, 300, 1400.0, 1600.0, 1e-2, 1e-4)
    assert len(ri) == 0 and torch.all(pix == 1.0)


def test_full_grid_constant_field_matches_beer_lambert_chord():
    o, d = _rays(W=8)
    g = _grid(32, True)
    logit = -4.0                                   # sigma = 0.018: transmittance stays above the early-stop threshold
    pix, (ri, ts, te) = pipeline.render_rays(lambda x: torch.full((len(x), 1), logit), g, ROI, o, d, 300, 1400.0, 1600.0, 1e-2, 1e-4)
    sig = 1.0 / (1.0 + np.exp(-logit))
    tmin, tmax = nerfacc_ref.ray_aabb_intersect(o, d, ROI, 1400.0, 1600.0)
    chord = tmax - tmin
    hit = chord > 0
    # sum of step widths differs from the exact chord by at most one step (2/3)
    assert np.all(np.abs(-np.log(pix.numpy()[hit]) / sig - chord[hit]) <= 2.0 / 3.0 + 1e-3)
    counts = np.bincount(ri, minlength=len(o))
    assert np.all(np.abs(counts[hit] - chord[hit] / (200.0 / 300.0)) <= 1.0)


def test_single_occupied_voxel_sample_count():
    res = 16
    g = _grid(res, False)
    g.binary[8, 8, 8] = True                       # voxel [0,12.5]^3
    o = np.array([[6.0, 6.0, 1500.0]], np.float32); d = np.array([[0.0, 0.0, -1.0]], np.float32)
    tmin, tmax = nerfacc_ref.ray_aabb_intersect(o, d, ROI, 1400.0, 1600.0)
    ri, ts, te, off = nerfacc_ref.march(o, d, tmin, tmax, ROI, res, g.binary, np.float32(200.0 / 300.0))
    assert abs(len(ri) - 12.5 / (200.0 / 300.0)) <= 1.0       # ceil(chord / dt) +- 1
    mid = 0.5 * (ts + te)
    z = 1500.0 - mid
    assert np.all((z >= -1e-3) & (z <= 12.5 + 1e-3))           # every sample midpoint lies inside the voxel
    assert np.allclose(te - ts, 200.0 / 300.0, atol=1e-4)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.floats(0.0, 1.0))
def test_march_structure_properties(seed, density):
    rng = np.random.default_rng(seed)
    res = 16
    binary = rng.random((res,) * 3) < density
    o, d = _rays(theta=float(rng.uniform(0, 360)), phi=float(rng.uniform(0, 180)), W=6)
    tmin, tmax = nerfacc_ref.ray_aabb_intersect(o, d, ROI, 1400.0, 1600.0)
    ri, ts, te, off = nerfacc_ref.march(o, d, tmin, tmax, ROI, res, binary, np.float32(200.0 / 300.0))
    assert off[0] == 0 and np.all(np.diff(off) >= 0) and off[-1] == len(ri)          # offsets monotone, sum(counts) = n
    assert np.all(np.diff(ri) >= 0)                                                  # sorted by ray
    for r in range(len(o)):
        s = slice(off[r], off[r + 1])
        assert np.all(np.diff(ts[s]) > 0) and np.all(te[s] > ts[s])                  # ascending, non-empty intervals
        assert np.all(ts[s] >= tmin[r] - 1e-3) and np.all(0.5 * (ts[s] + te[s]) < tmax[r])
    # every emitted midpoint sits in an occupied cell
    mid = (0.5 * (ts + te))[:, None]
    pts = o[ri] + mid * d[ri]
    g = nerfacc_ref.OccupancyGrid(ROI, res); g.binary = binary
    assert np.all(g.query_occ(pts.astype(np.float32)) == 1.0)


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_visibility_and_composite_properties(seed):
    rng = np.random.default_rng(seed)
    R = 40
    counts = rng.integers(0, 30, R)
    off = np.zeros(R + 1, np.int64); np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    alphas = rng.random(n).astype(np.float32) * 0.5
    keep = nerfacc_ref.visibility(off, alphas, 1e-2, 0.0)
    for r in range(R):
        k = keep[off[r]:off[r + 1]]
        if len(k):
            assert k[0]                                             # T_0 = 1 >= eps
            assert np.all(np.diff(k.astype(int)) <= 0)              # once invisible, always invisible (alpha in [0,1))
    # composite: permutation invariance within a ray and empty rays -> 1
    ri = np.repeat(np.arange(R), counts)
    a = np.exp(-rng.random(n)).astype(np.float32)
    p1 = nerfacc_ref.scatter_mul(a, ri, R)
    perm = np.concatenate([off[r] + rng.permutation(counts[r]) for r in range(R)]).astype(np.int64) if n else np.zeros(0, np.int64)
    p2 = nerfacc_ref.scatter_mul(a[perm], ri, R)
    assert np.allclose(p1, p2, rtol=1e-5)
    assert np.all(p1[counts == 0] == 1.0)


def test_grid_update_semantics():
    g = nerfacc_ref.OccupancyGrid(ROI, 8)
    rng = np.random.default_rng(0)
    # step not divisible by 16: nothing happens
    g.every_n_step(3, lambda x: np.ones(len(x), np.float32), rng=rng)
    assert g.occs.max() == 0 and not g.binary.any()
    # warm-up step: every cell evaluated once; threshold = min(mean, occ_thre)
    g.every_n_step(0, lambda x: (x[:, 0] > 0).astype(np.float32) * 0.5, occ_thre=1e-2, rng=rng)
    assert g.binary.sum() == 8 ** 3 // 2
    before = g.occs.copy()
    g.every_n_step(16, lambda x: np.zeros(len(x), np.float32), occ_thre=1e-2, rng=rng)
    assert np.all(g.occs <= before) and np.all(g.occs >= before * np.float32(0.95) - 1e-9)     # EMA decay of touched cells


def test_march_resumes_from_the_end_of_any_sample():
    """The property lazy marching rests on: after emitting a sample the marcher's state is a function of
    that sample's t_end alone, so marching a ray again with t_min = t_end of its k-th sample reproduces samples k+1 ... bit
    for bit -- head (first k0 samples) + tail (resumed) = the full march.  Checked on the oracle for sparse and dense grids."""
    rng = np.random.default_rng(5)
    for fill, k0 in ((0.15, 32), (0.5, 7), (1.0, 32), (0.03, 1)):
        g = nerfacc_ref.OccupancyGrid(ROI, 32)
        g.binary[:] = rng.random(g.binary.shape) < fill
        o, d = _rays(theta=40.0, phi=25.0, W=12)
        t_min, t_max = nerfacc_ref.ray_aabb_intersect(o, d, ROI, 1400.0, 1600.0)
        step = np.float32(200.0 / 300)
        ri, ts, te, off = nerfacc_ref.march(o, d, t_min, t_max, ROI, 32, g.binary, step)
        cnt = np.diff(off)
        long_rays = np.nonzero(cnt > k0)[0]
        assert len(long_rays) > 10
        t_resume = t_min.copy()
        t_resume[long_rays] = te[off[long_rays] + k0 - 1]
        sel = long_rays
        ri2, ts2, te2, off2 = nerfacc_ref.march(o[sel], d[sel], t_resume[sel], t_max[sel], ROI, 32, g.binary, step)
        assert np.array_equal(np.diff(off2), cnt[sel] - k0)
        tail = np.concatenate([np.arange(off[r] + k0, off[r + 1]) for r in sel])
        assert np.array_equal(ts2, ts[tail]) and np.array_equal(te2, te[tail])


def test_visibility_of_head_plus_tail_equals_full_filter():
    """Second half of the lazy-marching argument: filtering the first k0 samples of every ray, carrying the transmittance
    over and filtering the rest only for rays that are still >= early_stop_eps keeps exactly the samples the reference's
    filter keeps on the full ray (alpha in [0, 1]: the transmittance never recovers)."""
    rng = np.random.default_rng(11)
    R, k0, eps, thre = 200, 32, np.float32(1e-2), np.float32(1e-3)
    cnt = rng.integers(0, 120, R)
    off = np.zeros(R + 1, np.int64); np.cumsum(cnt, out=off[1:])
    alphas = (rng.random(off[-1]) ** 3).astype(np.float32) * np.float32(0.4)
    alphas[rng.random(off[-1]) < 0.1] = 0.0
    full = nerfacc_ref.visibility(off, alphas, eps, thre)
    keep = np.zeros_like(full)
    for r in range(R):
        a = alphas[off[r]:off[r + 1]]
        T = np.float32(1.0)
        for j in range(min(k0, len(a))):                      # head
            keep[off[r] + j] = (T >= eps) and (a[j] >= thre)
            T = np.float32(T * np.float32(np.float32(1.0) - a[j]))
        if len(a) > k0 and T >= eps:                          # tail: only if the ray is still transparent behind its head
            for j in range(k0, len(a)):
                keep[off[r] + j] = (T >= eps) and (a[j] >= thre)
                T = np.float32(T * np.float32(np.float32(1.0) - a[j]))
    assert np.array_equal(keep, full)


# ------------------------------------------------------------------------------------------------ closed form of the fp32 t-chain
def _linear_chain(t_lo, t_hi, dt):
    """numpy restatement of `linear_chain` (csrc/march.cu): the guard under which the marcher's t <- fl(t + dt) adds the SAME exact
    increment at every step.  Returns the increment or None."""
    f32 = np.float32
    a, b = np.array([t_lo, t_hi], f32).view(np.uint32)
    e = int(a) >> 23
    if e != (int(b) >> 23) or e == 0 or e >= 255 or e < 24:
        return None
    u = np.array([(e - 23) << 23], np.uint32).view(f32)[0]           # ulp of the binade
    s2 = f32(f32(f32(dt) / u) * f32(2.0))
    if not s2 < f32(8388608.0):
        return None
    if s2 == np.floor(s2) and np.fmod(s2, f32(2.0)) == 1.0:          # dt / ulp ends in .5: round-to-even would depend on t
        return None
    delta = f32(f32(f32(t_lo) + f32(dt)) - f32(t_lo))
    return delta if delta > 0 else None


@settings(max_examples=200, deadline=None)
@given(st.integers(24, 200), st.floats(0.0, 1.0), st.floats(1e-4, 0.2), st.integers(1, 4000))
def test_closed_form_t_chain_matches_serial_fp32(expo, frac, rel_dt, k):
    """The warp-per-ray march passes compute sample k of a run as (t1 + (k-1) D, t1 + k D) instead of replaying
    t0 <- t1, t1 <- fl(t0 + dt) k times.  Inside one fp32 binade and under the guard this is the same bit pattern: every value is
    a multiple of the binade's ulp, D = fl(t + dt) - t does not depend on t, and the products / sums are exact."""
    f32 = np.float32
    lo = f32(2.0) ** f32(expo - 127)
    t0 = f32(lo * f32(1.0 + 0.5 * frac))                              # somewhere in the lower half of the binade
    dt = f32(lo * f32(rel_dt) / f32(16.0))
    t_end = f32(t0 + f32(k + 2) * dt)
    D = _linear_chain(t0, t_end, dt)
    if D is None:
        return                                                         # outside the guard the kernels walk serially
    a, b = t0, f32(t0 + dt)                                            # serial walk
    for _ in range(k):
        a, b = b, f32(b + dt)
    t1 = f32(t0 + dt)
    ca = f32(t1 + f32(f32(k - 1) * D))
    cb = f32(t1 + f32(f32(k) * D))
    assert a.view(np.uint32) == ca.view(np.uint32) and b.view(np.uint32) == cb.view(np.uint32)


def test_closed_form_guard_rejects_round_to_even_ties_and_binade_crossings():
    f32 = np.float32
    u = f32(2.0) ** f32(10 - 23)                                       # ulp of [1024, 2048)
    assert _linear_chain(f32(1400.0), f32(1600.0), f32(2.0 / 3.0)) is not None      # the benchmark geometry
    assert _linear_chain(f32(1400.0), f32(1600.0), f32(100.5) * u) is None          # dt / ulp = 100.5: a tie at every step
    assert _linear_chain(f32(900.0), f32(1100.0), f32(2.0 / 3.0)) is None           # crosses 1024
    assert _linear_chain(f32(1400.0), f32(1600.0), f32(0.0)) is None                # no progress
    # the tie case really is t-dependent (that is why it is excluded): fl(t + 100.5 u) - t differs between even and odd t / u
    t_even, t_odd = f32(1400.0), f32(1400.0) + u
    d = f32(100.5) * u
    assert f32(f32(t_even + d) - t_even) != f32(f32(t_odd + d) - t_odd)


@settings(max_examples=200, deadline=None)
@given(st.floats(0.0, 1.0), st.floats(0.0, 40.0), st.sampled_from([2.0 / 3.0, 0.4, 0.2, 1.25]))
def test_closed_form_empty_space_skip_matches_serial_fp32(frac, gap, dt):
    """Empty-space skipping (`march_count_warp_kernel`): the serial marcher adds dt to the midpoint until it reaches the next voxel
    boundary; in one binade that is j = max(1, ceil((target - t) / D)) steps, computed exactly on multiples of the ulp."""
    f32 = np.float32
    dt = f32(dt)
    D = _linear_chain(f32(1400.0), f32(1600.0), dt)
    assert D is not None
    u = f32(2.0) ** f32(10 - 23)
    tm = f32(f32(1400.0) + f32(150.0 * frac))
    target = f32(tm + f32(gap))
    t = tm                                                             # serial: do { t += dt } while (t < target)
    while True:
        t = f32(t + dt)
        if not t < target:
            break
    diff = int(np.rint(f32(f32(target - tm) / u)))
    Di = int(np.rint(f32(D / u)))
    j = max(1, (diff + Di - 1) // Di)
    closed = f32(tm + f32(f32(j) * D))
    assert t.view(np.uint32) == closed.view(np.uint32)
