"""Synthetic angiography data: analytic vessel phantom -> attenuation volume -> cone-beam projections, plus the
device-resident ray pool that replaces the reference's CSV/pandas DataFrame of precomputed rays.

The reference's datasets cannot be shared (README.md:18) and its ``load_data`` is missing
(nerf/run_nerf_acc.py:82), so the benchmark / test inputs are generated here following the reference's
phantom pipeline (/root/reference/phantomdata/cttoray.py:189-262, helpers.py:17-18,192-224):
volume -> trilinear interpolation along each ray -> Beer-Lambert product (a CUDA kernel of its own:
csrc/projector.cu).  This is input preparation, not the hot path.
"""
import math

import numpy as np
import torch

from . import ops
from .geometry import source_matrix


# ------------------------------------------------------------------------------------------------ phantom
def default_capsules(seed=0):
    """A small coronary-like tree: (p0[3], p1[3], radius) capsules inside +-60."""
    rng = np.random.default_rng(seed)
    caps = []
    root = np.array([-45.0, -40.0, -30.0])
    trunk_dir = np.array([0.6, 0.55, 0.45])
    p = root
    for i in range(4):
        q = p + trunk_dir * 28 + rng.normal(0, 4, 3)
        caps.append((p.copy(), q.copy(), 6.0 - i * 0.9))
        # side branch
        b = q + rng.normal(0, 1, 3) * 6 + np.array([(-1) ** i * 22.0, 12.0 * ((i % 2) * 2 - 1), 10.0])
        caps.append((q.copy(), np.clip(b, -60, 60), 3.2 - 0.3 * i))
        p = q
    return [(np.clip(a, -60, 60), np.clip(b, -60, 60), float(r)) for a, b, r in caps]


def capsule_sdf(points, capsules):
    """Signed distance of points[N,3] to the union of capsules (negative inside)."""
    d = None
    for a, b, r in capsules:
        a_t = torch.as_tensor(a, dtype=points.dtype, device=points.device)
        b_t = torch.as_tensor(b, dtype=points.dtype, device=points.device)
        ab = b_t - a_t
        t = ((points - a_t) @ ab / (ab @ ab)).clamp(0, 1)
        dist = (points - (a_t + t[:, None] * ab)).norm(dim=-1) - r
        d = dist if d is None else torch.minimum(d, dist)
    return d


def rev_sigmoid(x, c1=1.0, c2=0.0):
    """/root/reference/phantomdata/helpers.py:17-18"""
    return 1.0 / (1.0 + torch.exp(c1 * (x - c2)))


# CT transfer function knots (HU -> attenuation) "used for ALL experiments" and its binary variant, /root/reference/phantomdata/helpers.py:33-70
_TF_X = (0.0, 753.0, 1585.85, 2332.9, 3306.18, 4000.0)
_TF_Y = (0.0, 0.0, 0.05, 0.0, 0.2, 0.4)
_TF_Y_BINARY = (0.0, 0.0, 0.0, 0.0, 0.2, 0.4)


def transfer_func_ct(vals, binary=False):
    """Piecewise-linear HU -> attenuation map of the reference's CT phantoms (helpers.py:33-70): 0 below 0 HU, linear between
    the knots (each segment evaluated as m*x + b with the reference's m and b, float64), 0.4 from 4000 HU on."""
    v = np.asarray(vals).astype("float64")
    ys = _TF_Y_BINARY if binary else _TF_Y
    out = np.empty_like(v)
    out[v < _TF_X[0]] = ys[0]
    for k in range(5):
        x1, x2, y1, y2 = _TF_X[k], _TF_X[k + 1], ys[k], ys[k + 1]
        sel = (v >= x1) & (v < x2)
        m = (y1 - y2) / (x1 - x2)
        b = (x1 * y2 - x2 * y1) / (x1 - x2)
        out[sel] = m * v[sel] + b
    out[v >= _TF_X[5]] = ys[5]
    return out


def get_weighted_img(img, sampling_strategy="segmentation"):
    """Sampling-weight image of one projection (helpers.py:226-247, the `distance_pixel_value` column): mask of the pixels
    darker than 1 ('segmentation'; 'frangi' needs scikit-image), min-max normalised, Euclidean distance transform,
    min-max normalised again, + 1e-10 so that every pixel can be drawn.  Host side (scipy), once per projection."""
    from scipy.ndimage import distance_transform_edt
    img = np.asarray(img)
    if sampling_strategy == "frangi":
        try:
            from skimage.filters import frangi
        except ImportError as e:
            raise NotImplementedError("sampling_strategy='frangi' needs scikit-image, which is not installed") from e
        mask = frangi(img, alpha=0.5, beta=0.5)
    else:
        mask = np.zeros(img.shape)
        mask[img < 1] = 1
    mask -= np.min(mask)
    mask /= np.max(mask)
    w = distance_transform_edt(mask)
    w -= np.min(w)
    w /= np.max(w)
    w += 1e-10
    return w


def make_volume(resolution=256, half_extent=100.0, kind="sdf", device="cuda", seed=0, hu_mu_scale=0.03):
    """Attenuation volume mu[res,res,res] (indexed [x,y,z]) on the +-half_extent lattice.

    kind 'sdf': mu = rev_sigmoid(sdf, c1=2)           (phantomdata/helpers.py:93 convention, vessels ~1, air 0)
    kind 'ct' : vessels + a soft-tissue ellipsoid, scaled to CT-like line integrals
    kind 'ct_hu': synthetic HU field through the reference's CT transfer function (configs 3-4)
    """
    caps = default_capsules(seed)
    lin = torch.linspace(-half_extent, half_extent, resolution, device=device)
    vol = torch.empty((resolution,) * 3, dtype=torch.float32, device=device)
    for ix in range(resolution):                      # slab by slab keeps the temporary small
        yy, zz = torch.meshgrid(lin, lin, indexing="ij")
        pts = torch.stack([torch.full_like(yy, float(lin[ix])), yy, zz], dim=-1).reshape(-1, 3)
        sdf = capsule_sdf(pts, caps)
        mu = rev_sigmoid(sdf, c1=2.0)
        if kind == "ct":
            tissue = ((pts / torch.tensor([85.0, 70.0, 90.0], device=device)).norm(dim=-1) < 1.0).float()
            mu = 0.12 * mu + 0.0015 * tissue
        elif kind == "ct_hu":
            # CT-derived phantom (BASELINE config 3): a synthetic Hounsfield field -- air 0, a soft-tissue ellipsoid at
            # ~1500 HU, contrast-filled vessels 3300 HU at the wall rising to 4000 HU on the centre line -- pushed through the
            # reference's piecewise-linear transfer function (phantomdata/helpers.py:33-70), then scaled by `hu_mu_scale` per world
            # unit so that line integrals over the +-100 box stay in the range of the reference's 100x100x(420 steps) CT geometry.
            tissue = ((pts / torch.tensor([85.0, 70.0, 90.0], device=device)).norm(dim=-1) < 1.0).double()
            core = (-sdf / 3.0).clamp(0.0, 1.0).double()
            hu = mu.double() * (3300.0 + 700.0 * core) + (1.0 - mu.double()) * 1500.0 * tissue
            mu = torch.from_numpy(transfer_func_ct(hu.cpu().numpy())).to(device=device, dtype=torch.float32) * hu_mu_scale
        elif kind == "sdf":
            mu = 0.08 * mu
        else:
            raise ValueError(kind)
        vol[ix] = mu.reshape(resolution, resolution)
    return vol


@torch.no_grad()
def project(volume, rays_o, rays_d, near, far, n_samples=300, half_extent=100.0, kind="ct"):
    """Ground-truth projector (/root/reference/phantomdata/helpers.py:192-224) on the GPU (angio_project_volume): trilinear
    lookups at evenly spaced depths (the reference's non-stratified depth values), I = prod exp(-mu * dist * |d|) ('ct',
    :208-211) or prod exp(-mu) ('sdf', :213-215).  The volume spans [-half_extent, half_extent]^3."""
    depths = torch.linspace(float(near), float(far), int(n_samples), device=rays_o.device, dtype=torch.float32)
    h = float(half_extent)
    return ops.project_volume(volume.contiguous().float(), np.array([-h, -h, -h, h, h, h], np.float32), rays_o.contiguous().float(),
                              rays_d.contiguous().float(), depths, kind)


# ------------------------------------------------------------------------------------------------ ray pool
class RayPool:
    """All rays of all views, device-resident, in the compact form (cam2world[V,4,4] f64, pixels[V,H,W] f32,
    weights[V,H,W] f32).  Replaces the reference's DataFrame of precomputed, CSV-serialised rays
    (nerf/run_nerf_acc.py:113-117): a ray is the integer triple (view, x, y) and is expanded by the ray-generation
    kernel when sampled."""

    def __init__(self, cam2world, pixels, focal, weights=None, train_on_test_view=True):
        self.cam2world = cam2world.contiguous()                 # [V,4,4] float64, CUDA
        self.pixels = pixels.contiguous()                       # [V,H,W] float32
        self.focal = float(focal)
        self.n_views, self.img_h, self.img_w = pixels.shape
        self.weights = None if weights is None else weights.contiguous().float()
        self.n_rays = self.n_views * self.img_h * self.img_w
        # The reference trains on the test view too (`train_ray_df = ray_df.copy()`, run_nerf_acc.py:114; the test view is the
        # LAST one, :85).  train_on_test_view=False draws training rays from the other views only.
        self.train_on_test_view = bool(train_on_test_view)
        self.n_train_rays = self.n_rays if self.train_on_test_view else (self.n_views - 1) * self.img_h * self.img_w
        self._wsum = None
        self._seed_streams = {}
        self._bufs = ops.BufferPool()
        self.last_status = None       # device int32[2] of the last draw: [candidates, overflow flag]

    @property
    def device(self):
        return self.pixels.device

    def rays_of_view(self, v):
        """(o[H*W,3], d[H*W,3], pix[H*W]) of one view, pixel order row-major [y, x]."""
        o, d = ops.raygen(self.cam2world, self.img_w, self.img_h, self.focal, view=int(v))
        return o, d, self.pixels[int(v)].reshape(-1)

    @torch.no_grad()
    def sample_ids(self, n, weights=None, generator=None, status=None):
        """Weighted sampling without replacement of n ray ids (exponential-race / Efraimidis-Spirakis keys with a
        threshold pre-filter so only ~n + 8 sqrt(n) candidates reach the exact selection), then a random shuffle -- the
        reference's ``DataFrame.sample(n, weights).sample(frac=1)`` (nerf/nerf_helpers.py:139).  Two kernel launches, no sync."""
        N, dev = self.n_train_rays, self.device
        if n > N:
            raise ValueError("cannot sample more rays than the pool holds without replacement")
        w = self.weights if weights is None else (None if isinstance(weights, str) and weights == "uniform" else weights)
        if w is not None:
            wf = w.reshape(-1).contiguous().float()[:N]          # ray ids are view-major: a prefix = the training views
            if weights is not None and not isinstance(weights, str):
                wsum, wsum2, npos = float(wf.sum().item()), float((wf.double() ** 2).sum().item()), int((wf > 0).sum().item())
            else:
                if self._wsum is None:                           # one-off reductions when the weight image is first used
                    self._wsum = (float(wf.sum().item()), float((wf.double() ** 2).sum().item()), int((wf > 0).sum().item()))
                wsum, wsum2, npos = self._wsum
            if n > npos:     # numpy / pandas: "Fewer non-zero entries in p than size" -- the draw could never be filled
                raise ValueError(f"cannot sample {n} rays without replacement: only {npos} rays have a positive weight")
        else:
            wf, wsum, wsum2 = None, float(N), float(N)
        # per-call 62-bit seed from a host-side stream tied to the torch generator's seed: reproducible, rank-dependent,
        # and no device->host sync
        key = id(generator) if generator is not None else None
        rng = self._seed_streams.get(key)
        if rng is None:
            rng = np.random.default_rng(generator.initial_seed() if generator is not None else None)
            self._seed_streams[key] = rng
        seed = int(rng.integers(0, 2 ** 62))
        ids, self.last_status = ops.sample_without_replacement(n, N, wf, wsum, wsum2, seed, dev, pool=self._bufs, status=status)
        return ids

    def gather(self, ids):
        return ops.raygen_flat(self.cam2world, ids, self.img_w, self.img_h, self.focal, pixels=self.pixels)

    def sample(self, n, weights=None, generator=None, status=None):
        """status: optional int32[2] device tensor receiving [candidates, overflow flag] (see ops.sample_without_replacement)."""
        return self.gather(self.sample_ids(n, weights=weights, generator=generator, status=status))


class ExplicitRayPool:
    """Device-resident copy of a reference ray DataFrame (one row per ray with precomputed origins / directions, the layout
    `run_nerf_acc.py:113-117` samples from): o[N,3], d[N,3], pixel_value[N] and any number of named weight columns.  Lets
    `sample_pixel_rays(train_ray_df, ...)` keep the reference's DataFrame signature while the draw itself runs in the sampler
    kernels; prefer `RayPool` (rays regenerated from (view, x, y)) when the projection matrices are known."""

    def __init__(self, rays_o, rays_d, pixel_values, weights=None):
        self.o = rays_o.contiguous().float()
        self.d = rays_d.contiguous().float()
        self.pix = pixel_values.contiguous().float()
        self.weight_columns = {k: v.contiguous().float() for k, v in (weights or {}).items()}
        self.n_rays = self.o.shape[0]
        self._wsum, self._seed_streams, self._bufs, self.last_status = {}, {}, ops.BufferPool(), None

    @classmethod
    def from_dataframe(cls, ray_df, device="cuda", weight_columns=("distance_pixel_value",)):
        def col3(prefix):
            if f"{prefix}_x" in ray_df.columns:
                a = np.stack([ray_df[f"{prefix}_{c}"].to_numpy(dtype=np.float64) for c in "xyz"], axis=1)
            else:
                a = np.asarray(ray_df[prefix].tolist(), dtype=np.float64)
            return torch.from_numpy(a.astype(np.float32)).to(device)      # the reference casts with .float() / torch.Tensor(...)
        w = {c: torch.from_numpy(ray_df[c].to_numpy(dtype=np.float64).astype(np.float32)).to(device) for c in weight_columns if c in ray_df.columns}
        pix = torch.from_numpy(ray_df["pixel_value"].to_numpy(dtype=np.float64).astype(np.float32)).to(device)
        return cls(col3("ray_origins"), col3("ray_directions"), pix, w)

    @property
    def device(self):
        return self.o.device

    @torch.no_grad()
    def sample(self, n, weights=None, generator=None, status=None):
        """weights: None (uniform) or the name of a weight column, like DataFrame.sample(n, weights=<column>)."""
        if n > self.n_rays:
            raise ValueError("cannot sample more rays than the pool holds without replacement")
        wf, wsum, wsum2 = None, float(self.n_rays), float(self.n_rays)
        if weights is not None:
            wf = self.weight_columns[weights]
            if weights not in self._wsum:
                self._wsum[weights] = (float(wf.sum().item()), float((wf.double() ** 2).sum().item()), int((wf > 0).sum().item()))
            wsum, wsum2, npos = self._wsum[weights]
            if n > npos:
                raise ValueError(f"cannot sample {n} rays without replacement: only {npos} rays have a positive weight")
        key = id(generator) if generator is not None else None
        rng = self._seed_streams.get(key)
        if rng is None:
            rng = np.random.default_rng(generator.initial_seed() if generator is not None else None)
            self._seed_streams[key] = rng
        ids, self.last_status = ops.sample_without_replacement(n, self.n_rays, wf, wsum, wsum2, int(rng.integers(0, 2 ** 62)), self.device,
                                                               pool=self._bufs, status=status)
        return self.o[ids], self.d[ids], self.pix[ids]


def make_dataset(img_size=64, thetas=(0.0, 45.0, 90.0, 135.0), test_view=(135.0, 135.0), kind="ct", volume_res=128,
                 n_proj_samples=300, src_dist=1500.0, half_extent=100.0, device="cuda", seed=0, weight_strategy="random",
                 train_on_test_view=True):
    """Views theta in `thetas` (phi = 0) + the test view (theta, phi) LAST (the reference treats the last projection
    as the test view and keeps it in the training pool, nerf/run_nerf_acc.py:85,114).  Focal length 7.5*W makes the
    detector span the AABB at the isocentre.  Returns (RayPool, info dict)."""
    focal = 7.5 * img_size
    src = np.array([0.0, 0.0, src_dist])
    views = [(float(t), 0.0) for t in thetas] + [tuple(map(float, test_view))]
    mats = np.stack([source_matrix(src, th, ph, 0) for th, ph in views])
    cam = torch.from_numpy(mats).to(device)
    vol = make_volume(volume_res, half_extent, kind=kind, device=device, seed=seed)
    near, far = src_dist - half_extent * 1.8, src_dist + half_extent * 1.8
    pix = torch.empty((len(views), img_size, img_size), dtype=torch.float32, device=device)
    for v in range(len(views)):
        o, d = ops.raygen(cam, img_size, img_size, focal, view=v)
        img = project(vol, o, d, near, far, n_proj_samples, half_extent, kind="sdf" if kind == "sdf" else "ct")
        if kind == "sdf":                                        # per-image min-max normalisation (sdftoray.py:125-126)
            img = (img - img.min()) / (img.max() - img.min() + 1e-12)
        pix[v] = img.view(img_size, img_size)
    if weight_strategy == "random":                              # cttoray.py:220-221: all ones
        weights = None
    elif weight_strategy == "segmentation":                      # helpers.py:229-244 without the EDT: vessel pixels up-weighted
        weights = (pix < pix.flatten(1).mean(dim=1)[:, None, None]).float() + 1e-3
    elif weight_strategy == "distance":                          # the reference's weight image (mask -> EDT), per projection on the host
        weights = torch.stack([torch.from_numpy(get_weighted_img(p.cpu().numpy())).float() for p in pix]).to(device)
    else:
        raise ValueError(weight_strategy)
    info = dict(focal=focal, src_dist=src_dist, near=src_dist - half_extent, far=src_dist + half_extent, views=views,
                half_extent=half_extent, volume=vol)
    return RayPool(cam, pix, focal, weights, train_on_test_view=train_on_test_view), info
