"""The slice of the third-party ``nerfacc`` 0.3.x API that the reference imports
(/root/reference/nerf/run_nerf_acc.py:12,197-198; nerf/nerf_helpers_acc.py:29,72-76;
visualization/visualization.py:162,214), implemented on the sm_100a kernels of libangio_b200.so.

    from nerf_for_angiography_b200.nerfacc import OccupancyGrid, ContractionType, ray_marching

Semantics follow nerfacc 0.3.5 as restated in oracle/nerfacc_ref.py + oracle/march_ref.c (the library is not
vendored by the reference and not installable here; see DESIGN.md "Oracle").
"""
import enum

import torch

from . import ops


class ContractionType(enum.Enum):
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2


class OccupancyGrid(torch.nn.Module):
    """Occupancy grid with an axis-aligned ROI: ``occs`` fp32 [res^3] (EMA of the occupancy field) and
    ``binary`` bool [res,res,res].  ``every_n_step`` refreshes it every ``n`` steps exactly like nerfacc's
    ``OccupancyGrid._update`` (all cells during warm-up, then N/4 uniform + N/4 occupied cells)."""

    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution=128, contraction_type=ContractionType.AABB):
        super().__init__()
        if contraction_type != ContractionType.AABB:
            raise NotImplementedError("only ContractionType.AABB is used by the reference and implemented")
        if isinstance(resolution, (list, tuple)):
            if len(set(resolution)) != 1:
                raise NotImplementedError("anisotropic grid resolution")
            resolution = resolution[0]
        self._resolution = int(resolution)
        self.num_cells = self._resolution ** 3
        self._contraction_type = contraction_type
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32).reshape(6).clone()
        self.register_buffer("_roi_aabb", roi)
        self.register_buffer("occs", torch.zeros(self.num_cells, dtype=torch.float32))
        self.register_buffer("_binary", torch.zeros((self._resolution,) * 3, dtype=torch.bool))
        self._roi_host = roi.cpu().numpy().copy()
        # mean(occs) caps alpha_thre in nerfacc.ray_marching.  It lives on the device (the visibility kernels read it there, so a
        # training step never waits for it) and is tied to occs._version: occs restored through load_state_dict / copy_ / fill_
        # invalidates it, and the next reader recomputes it (the kernels of _update write through raw pointers and record the
        # version themselves).
        self.register_buffer("_occs_mean", torch.zeros(1, dtype=torch.float32), persistent=False)
        self._mean_version = None
        self._mean_host = None

    def occs_mean_dev(self):
        """float32[1] device tensor holding mean(occs) (no host synchronisation)."""
        if self._mean_version != self.occs._version:
            torch.mean(self.occs, dim=0, keepdim=True, out=self._occs_mean)
            self._mean_version, self._mean_host = self.occs._version, None
        return self._occs_mean

    @property
    def occs_mean_host(self):
        """mean(occs) as a python float (one device read when it is not cached) -- nerfacc evaluates `grid.occs.mean().item()`
        on every ray_marching call."""
        m = self.occs_mean_dev()
        if self._mean_host is None:
            self._mean_host = float(m.item())
        return self._mean_host

    @occs_mean_host.setter
    def occs_mean_host(self, value):
        self._occs_mean.fill_(float(value))
        self._mean_version, self._mean_host = self.occs._version, float(value)

    # nerfacc properties
    @property
    def roi_aabb(self):
        return self._roi_aabb

    @property
    def binary(self):
        return self._binary

    @property
    def resolution(self):
        return torch.tensor([self._resolution] * 3, dtype=torch.int32)

    @property
    def contraction_type(self):
        return self._contraction_type

    @property
    def device(self):
        return self.occs.device

    def _binary_u8(self):
        b = self._binary
        if b.dtype != torch.bool or not b.is_contiguous() or b.numel() != self.num_cells:
            # `_binary` may have been assigned from outside (visualization.py:162)
            b = b.to(device=self.occs.device, dtype=torch.bool).contiguous().reshape((self._resolution,) * 3)
            self._binary = b
        return b.view(torch.uint8)

    @torch.no_grad()
    def _sample_cells(self, step, warmup_steps, generator=None):
        if step < warmup_steps:
            return None                                   # all cells
        n = self.num_cells // 4
        dev = self.device
        uniform = torch.randint(self.num_cells, (n,), device=dev, generator=generator)
        occupied = torch.nonzero(self._binary.flatten())[:, 0]
        if n < len(occupied):
            sel = torch.randint(len(occupied), (n,), device=dev, generator=generator)
            occupied = occupied[sel]
        return torch.cat([uniform, occupied], dim=0)

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre=0.01, ema_decay=0.95, warmup_steps=256, cells=None, jitter=None,
                generator=None):
        if cells is None:
            cells = self._sample_cells(step, warmup_steps, generator)
        n = self.num_cells if cells is None else cells.numel()
        if jitter is None:
            jitter = torch.rand((n, 3), dtype=torch.float32, device=self.device, generator=generator)
        x = ops.grid_cell_points(cells, jitter.contiguous(), self._roi_host, self._resolution)
        occ = occ_eval_fn(x)
        occ = occ.reshape(-1).contiguous().float()
        ops.grid_ema_update(self.occs, cells, occ, ema_decay)
        mean = ops.grid_threshold(self.occs, occ_thre, self._binary_u8())
        self._occs_mean.copy_(mean)                         # stays on the device: no host sync in a grid refresh
        self._mean_version, self._mean_host = self.occs._version, None

    @torch.no_grad()
    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16, **kw):
        if not self.training:
            raise RuntimeError("You should only call this function only during training. Please call _update() directly "
                               "if you want to update the field during inference.")
        if step % n == 0 and self.training:
            self._update(step=step, occ_eval_fn=occ_eval_fn, occ_thre=occ_thre, ema_decay=ema_decay,
                         warmup_steps=warmup_steps, **kw)

    @torch.no_grad()
    def query_occ(self, samples):
        pts = samples.reshape(-1, 3).contiguous().float()
        return ops.grid_query(pts, self._roi_host, self._resolution, self._binary_u8()).reshape(samples.shape[:-1])

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._roi_host = self._roi_aabb.detach().cpu().numpy().copy()
        self._mean_version = None
        return out


def _is_fused_model(m):
    from .model.CPPN import CPPN
    return isinstance(m, CPPN)


class PackedRayIndices(torch.Tensor):
    """nerfacc's `ray_indices` (int64 ray id per sample, sorted by ray) that also carries the packed layout it came from: the
    segment offsets [n_rays + 1] int32 and the int32 index vector the kernels read.  `acc_render_volume_density` and the fused
    model queries take them from here instead of rebuilding the offsets with a search (and a synchronising sortedness check).

    Copies that keep every sample in place (`clone`, `detach`, `contiguous`, `long`, same-dtype / same-device `to`) keep the layout;
    anything else -- slicing, masking, arithmetic, using it as an index -- returns a plain tensor."""

    _KEEP = ("clone", "detach", "contiguous", "long", "to", "cuda")

    @staticmethod
    def wrap(ray_indices, offsets, idx32):
        t = ray_indices.as_subclass(PackedRayIndices)
        t._angio_offsets = offsets
        t._angio_idx32 = idx32
        return t

    def __deepcopy__(self, memo):
        off, i32 = getattr(self, "_angio_offsets", None), getattr(self, "_angio_idx32", None)
        plain = self.as_subclass(torch.Tensor).clone()
        return plain if off is None else PackedRayIndices.wrap(plain, off.clone(), i32.clone())

    def __reduce_ex__(self, proto):                    # pickles / torch.save as the plain index tensor
        return self.as_subclass(torch.Tensor).__reduce_ex__(proto)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **kwargs)
        src = args[0] if args and isinstance(args[0], cls) else None
        if (src is not None and isinstance(out, torch.Tensor) and getattr(func, "__name__", "") in cls._KEEP
                and out.shape == src.shape and out.dtype == src.dtype and out.device == src.device
                and getattr(src, "_angio_offsets", None) is not None):
            return cls.wrap(out.as_subclass(torch.Tensor), src._angio_offsets, src._angio_idx32)
        if isinstance(out, cls):                       # never leak the subclass (and stale offsets) into derived tensors
            out = out.as_subclass(torch.Tensor)
        return out


@torch.no_grad()
def ray_marching(rays_o, rays_d, t_min=None, t_max=None, scene_aabb=None, grid=None, sigma_fn=None, alpha_fn=None,
                 early_stop_eps=1e-4, alpha_thre=0.0, near_plane=None, far_plane=None, render_step_size=1e-3,
                 stratified=False, cone_angle=0.0, radiance_field=None, return_offsets=False):
    """nerfacc.ray_marching(...) for the argument combination the reference uses
    (/root/reference/nerf/nerf_helpers_acc.py:29): scene_aabb + grid + alpha_fn + near/far planes.

    ``radiance_field`` (extension): a CPPN whose alpha = 1 - exp(-sigmoid(f(x)) * dt) is evaluated by the fused
    MLP kernel straight from (ray, t0, t1) -- equivalent to the reference's alpha_fn closure without
    materialising positions.  Returns (ray_indices int64 [n], t_starts [n,1], t_ends [n,1]).
    """
    if t_min is not None or t_max is not None or stratified or cone_angle != 0.0 or sigma_fn is not None:
        raise NotImplementedError("ray_marching: only the reference's call pattern is implemented "
                                  "(scene_aabb, grid, alpha_fn, near/far planes, fixed step)")
    if grid is None or scene_aabb is None:
        raise NotImplementedError("ray_marching: grid and scene_aabb are required")
    rays_o = rays_o.contiguous().float()
    rays_d = rays_d.contiguous().float()
    near = -1e10 if near_plane is None else float(near_plane)
    far = 1e10 if far_plane is None else float(far_plane)
    ray_idx, t0, t1, offsets = ops.march(rays_o, rays_d, scene_aabb, grid._roi_host, grid._resolution, grid._binary_u8(),
                                         near, far, render_step_size)
    have_fn = alpha_fn is not None or radiance_field is not None
    if (alpha_thre > 0.0 or early_stop_eps > 0.0) and have_fn and ray_idx.numel() > 0:
        if radiance_field is not None and _is_fused_model(radiance_field):
            m = radiance_field
            if early_stop_eps > 0.0:
                # two-phase visibility pass with early ray termination: same kept samples as evaluating every sample
                m._ensure_flat()
                kp = m._kernel_params()
                packed = m._packed_weights(kp) if m._precision_id == ops.PREC_BF16 else None
                alphas, _ = ops.alphas_two_phase(m._desc, kp, packed, m._precision_id, rays_o, rays_d, ray_idx, t0, t1,
                                                 offsets, early_stop_eps)
            else:
                alphas = m.query(ops.OUT_ALPHA, rays_o=rays_o, rays_d=rays_d, ray_idx=ray_idx, t_starts=t0, t_ends=t1)
        else:
            fn = alpha_fn
            alphas = fn(t0[:, None], t1[:, None], ray_idx.long()).reshape(-1).contiguous().float()
        thre = min(float(alpha_thre), float(grid.occs_mean_host))
        ray_idx, t0, t1, offsets, _ = ops.visibility_compact(alphas, offsets, t0, t1, early_stop_eps, thre)
    ray_indices = PackedRayIndices.wrap(ray_idx.long(), offsets, ray_idx)   # the packed layout rides along for the compositor
    out = (ray_indices, t0[:, None], t1[:, None])
    return out + (offsets,) if return_offsets else out
