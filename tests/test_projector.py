"""Ground-truth projector: the oracle's trilinear interpolation is pinned against scipy's
RegularGridInterpolator -- the library the reference's phantom scripts call -- and the CUDA projector against the oracle."""
import numpy as np
import pytest
import torch

from oracle import geometry as ogeo, projector as oproj


def _volume(n=(20, 24, 28), seed=0):
    rng = np.random.default_rng(seed)
    v = rng.random(n) * 0.02
    v[:, :, -1] = 0.0          # empty far face: the reference's 1e10 tail multiplies exactly 0 there
    return v


def test_oracle_trilinear_matches_scipy():
    from scipy.interpolate import RegularGridInterpolator
    v = _volume()
    lo, hi = np.array([-100.0, -80.0, -60.0]), np.array([100.0, 80.0, 60.0])
    axes = [np.linspace(lo[k], hi[k], v.shape[k]) for k in range(3)]
    interp = RegularGridInterpolator(axes, v, method="linear", bounds_error=False, fill_value=0)
    rng = np.random.default_rng(1)
    pts = rng.uniform(-120, 120, size=(5000, 3))
    pts[:50] = np.stack(np.meshgrid(*[[lo[k], hi[k]] for k in range(3)], indexing="ij"), -1).reshape(-1, 3).repeat(7, 0)[:50]   # corners / faces
    assert np.allclose(oproj.trilinear(v, lo, hi, pts), interp(pts), rtol=1e-12, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ct", "sdf"])
def test_cuda_projector_vs_oracle(kind):
    import nerf_for_angiography_b200 as A
    v = _volume((40, 40, 40), seed=2)
    lo, hi = np.full(3, -100.0), np.full(3, 100.0)
    os_, ds_ = [], []
    for th, ph in ((0.0, 0.0), (50.0, 0.0), (135.0, 135.0)):
        o, d, _ = ogeo.get_ray_values(th, ph, 0.0, [0, 0, 1500.0], 24, 20, 7.5 * 24)
        os_.append(o.reshape(-1, 3)); ds_.append(d.reshape(-1, 3))
    o = np.concatenate(os_); d = np.concatenate(ds_)
    o[:40] += np.array([400.0, 0.0, 0.0])                             # forty rays that miss the volume: pixel exactly 1
    depths = np.linspace(1320.0, 1680.0, 300)
    ref = oproj.ray_tracing(v, lo, hi, o, d, depths, kind)
    got = A.ops.project_volume(torch.from_numpy(v).float().cuda(), np.concatenate([lo, hi]).astype(np.float32),
                               torch.from_numpy(o).float().cuda().contiguous(), torch.from_numpy(d).float().cuda().contiguous(),
                               torch.from_numpy(depths).float().cuda(), kind).cpu().numpy()
    assert ref.min() < 0.9 and ref.max() > 0.999                      # rays through the volume and rays that miss it
    assert np.abs(got - ref).max() <= 2e-4, np.abs(got - ref).max()   # fp32 positions / interpolation vs the float64 oracle
