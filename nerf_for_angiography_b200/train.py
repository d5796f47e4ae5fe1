"""Training / evaluation loop of the reference driver (/root/reference/nerf/run_nerf_acc.py:263-328, 338-357) on the
fused kernels, single- or multi-GPU (one process per GPU, rays sharded, gradients all-reduced with NCCL).

One ``Trainer.step()`` is one reference iteration:
  sample_pixel_rays -> acc_update_n_step (x2 grids) -> acc_ray_marching (march + no-grad MLP + visibility filter)
  -> MLP forward on the kept samples -> Beer-Lambert composite + MSE -> backward -> Adam -> lr decay.
Autograd is not involved: the composite/MSE tail and the MLP backward are explicit kernels on flat buffers.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib, ops
from .model.CPPN import CPPN
from .nerfacc import ContractionType, OccupancyGrid


class StepResult(dict):
    """Result of Trainer.step().  `loss` is a device scalar; the sample counts live in `totals` (device int32
    [marched, kept, sampler candidates, sampler overflow]) and are fetched -- one D2H copy -- only when
    `n_samples_prefilter` / `n_samples` are actually read, so a training loop that does not look at them never syncs."""

    def _resolve(self):
        if "_host_totals" not in self:
            t = dict.__getitem__(self, "totals")
            dict.__setitem__(self, "_host_totals", t.tolist() if isinstance(t, torch.Tensor) else list(t))
        return dict.__getitem__(self, "_host_totals")

    def __getitem__(self, key):
        if key in ("n_samples_prefilter", "n_samples") and not dict.__contains__(self, key):
            h = self._resolve()
            if h[3] != 0:
                raise RuntimeError("ray sampler: candidate buffer overflow / underflow -- re-draw with a larger threshold")
            return h[0] if key == "n_samples_prefilter" else h[1]
        return dict.__getitem__(self, key)


class Trainer:
    def __init__(self, model: CPPN, pool, near, far, n_rays=5625, n_steps=300, half_extent=100.0, grid_resolution=128,
                 lr=1e-4, decay_rate=0.1, decay_steps=500000, early_stop_eps=1e-2, alpha_thre=1e-4, vessel_alpha_thre=5e-2,
                 vessel_grid=True, seed=0, process_group=None, sync_free="auto", memory_fraction=0.6, early_termination=32):
        self.model = model
        self.pool = pool
        self.dev = pool.device
        self.near, self.far = float(near), float(far)
        self.n_rays, self.n_steps = int(n_rays), int(n_steps)
        self.step_size = (self.far - self.near) / self.n_steps            # nerf_helpers_acc.py:27
        self.lr0, self.decay_rate, self.decay_steps = lr, decay_rate, decay_steps
        self.early_stop_eps, self.alpha_thre, self.vessel_alpha_thre = early_stop_eps, alpha_thre, vessel_alpha_thre
        o = half_extent
        self.scene_aabb = torch.tensor([-o, -o, -o, o, o, o], dtype=torch.float32, device=self.dev)   # run_nerf_acc.py:196
        self._aabb_host = np.array([-o, -o, -o, o, o, o], dtype=np.float32)                           # host copy: no D2H per step
        self.acc_grid = OccupancyGrid(self.scene_aabb, grid_resolution, ContractionType.AABB).to(self.dev)
        self.vessel_acc_grid = OccupancyGrid(self.scene_aabb, grid_resolution, ContractionType.AABB).to(self.dev) if vessel_grid else None
        model._ensure_flat()
        self.flat = model._flat
        self.kflat = self.flat
        # one extra slot behind the gradient carries the kept-sample count through the all-reduce (Adam skips on 0)
        self.grad = torch.zeros(self.flat.numel() + 1, dtype=torch.float32, device=self.dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.packed = None
        self.pool_bufs = ops.BufferPool()
        self.n_iter = 0
        self.lr = lr
        self.pg = process_group
        self.world = 1
        self.rank = 0
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        # gradients of all ranks in NVLink peer memory -> all-reduce fused into the Adam kernel (ANGIO_P2P=0: NCCL all_reduce)
        self.peer = None
        if self.world > 1 and os.environ.get("ANGIO_P2P", "1") != "0" and model._precision_id == ops.PREC_BF16:
            try:
                from .distributed import PeerGradients
                self.peer = PeerGradients(self.flat.numel() + 1, self.dev, process_group)
            except Exception as e:                                       # no symmetric memory on this system: NCCL path
                if os.environ.get("ANGIO_P2P") == "1":
                    raise
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable ({e!r}); using NCCL all_reduce")
        # rays: every rank draws a different batch; grids: every rank draws the SAME cells/jitter
        self.ray_gen = torch.Generator(device=self.dev).manual_seed(seed + 1000 * self.rank + 1)
        self.grid_gen = torch.Generator(device=self.dev).manual_seed(seed)
        self.last = {}
        self.kernel_events = None     # bench.py: list of (start event, end event, sample count) per visibility-pass MLP launch
        self.sync_free = self._plan_memory(sync_free, memory_fraction)
        if self.world > 1:
            # the peer-memory gradient exchange needs every rank on the same path: sync-free only if ALL ranks can afford it
            flag = torch.tensor([1 if self.sync_free else 0], dtype=torch.int32, device=self.dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=process_group)
            self.sync_free = bool(flag.item())
        # draw + march the next batch under the gradient all-reduce (sync-free mode); ANGIO_PREFETCH=0 keeps the plain order
        self.prefetch = os.environ.get("ANGIO_PREFETCH", "1") != "0"
        self.shard_grid = os.environ.get("ANGIO_SHARD_GRID", "1") != "0"
        self.train_memory_bytes = int(os.environ.get("ANGIO_TRAIN_MEMORY_GB", "0")) * 2 ** 30 or int(0.4 * torch.cuda.mem_get_info(self.dev)[0])
        self._prefetched = None
        self._march_calls = 0
        # visibility pass with early ray termination (bf16 path): number of leading samples per ray evaluated before the rays
        # that are already opaque are dropped (multiple of 32); 0 / None = evaluate every marched sample like the reference
        self.early_termination = int(early_termination) if early_termination else 0
        # lazy marching (sync-free loop with early termination): march only the first samples of every ray before the
        # visibility pass and the rest only for rays that are still transparent behind them.  ANGIO_LAZY_MARCH=0: full march.
        self.lazy_march = bool(self.sync_free and self.early_termination and os.environ.get("ANGIO_LAZY_MARCH", "1") != "0")
        # sampler status of past draws, copied to pinned host memory without waiting and inspected once the copy has landed
        self._status_host = torch.zeros(2, dtype=torch.int32).pin_memory() if self.dev.type == "cuda" else None
        self._status_event = None

    # ------------------------------------------------------------------ memory plan
    def _plan_memory(self, mode, fraction):
        """The sync-free loop sizes every per-sample buffer for the worst case (each ray keeps every march step) so that no
        sample count ever has to reach the host: 12 B/sample marched + 12 B/sample kept + alphas / logits / gradients / mask
        + the bf16 tile images of the training forward/backward.  Used when that fits in `fraction` of the free HBM (config 3:
        ~55 GB of 180 GB); otherwise the loop reads the kept count once per step and grows its buffers geometrically."""
        if mode is False or self.model._precision_id != ops.PREC_BF16:
            return False
        lib, desc, prec = _lib.load(), self.model._desc, self.model._precision_id
        cap = ops.march_capacity(self.n_rays, self.near, self.far, self.step_size)
        need = cap * 48 + int(lib.angio_mlp_saved_bytes(ctypes.byref(desc), cap, prec)) + \
            int(lib.angio_mlp_workspace_bytes(ctypes.byref(desc), cap, prec, 1))
        free, _ = torch.cuda.mem_get_info(self.dev)
        if need > fraction * free:
            if mode is True:
                raise RuntimeError(f"sync-free training needs {need / 2**30:.1f} GiB of buffers, {free / 2**30:.1f} GiB are free")
            return False
        self.capacity = cap
        self.pool_bufs.reserve("saved", int(lib.angio_mlp_saved_bytes(ctypes.byref(desc), cap, prec)), self.dev)
        self.pool_bufs.reserve("bwd_ws", int(lib.angio_mlp_workspace_bytes(ctypes.byref(desc), cap, prec, 1)), self.dev)
        return True

    # ------------------------------------------------------------------ pieces
    def _occ_eval(self, x):
        """occ_eval_fn of acc_update_n_step (sigma at the jittered cell points).  Data-parallel runs SHARD the refresh: every rank
        draws the same cells / jitter (same generator), evaluates a contiguous 1/world slice of them and all-gathers the
        occupancies (the per-sample result of the kernel does not depend on which rank or tile computed it, so every rank
        ends up with bit-identical grids).  ANGIO_SHARD_GRID=0: every rank evaluates every cell."""
        n = x.shape[0]
        if self.world == 1 or not self.shard_grid or n < self.world:
            return ops.mlp_forward(self.model._desc, self.kflat, self.packed, ops.OUT_SIGMA, self.model._precision_id, points=x)
        from .distributed import evaluate_sharded
        chunk = (n + self.world - 1) // self.world
        return evaluate_sharded(lambda xs: ops.mlp_forward(self.model._desc, self.kflat, self.packed, ops.OUT_SIGMA, self.model._precision_id,
                                                           points=xs),
                                x, self.rank, self.world, group=self.pg,
                                out=self.pool_bufs.typed("occ_gather", chunk * self.world, torch.float32, self.dev))

    def update_grids(self):
        """acc_update_n_step for both grids (run_nerf_acc.py:285-286)."""
        self.acc_grid.every_n_step(self.n_iter, self._occ_eval, occ_thre=self.alpha_thre, n=self.GRID_EVERY, generator=self.grid_gen)
        if self.vessel_acc_grid is not None:
            self.vessel_acc_grid.every_n_step(self.n_iter, self._occ_eval, occ_thre=self.vessel_alpha_thre, n=self.GRID_EVERY,
                                              generator=self.grid_gen)

    GRID_EVERY = 16     # nerfacc's every_n_step default (the reference does not override it)

    def _march(self, o, d, totals, pooled=True):
        """Marcher only (bf16 path: capacity-sized arrays, count left on the device in offsets[R] and totals[0])."""
        g = self.acc_grid
        bf16 = self.model._precision_id == ops.PREC_BF16
        if pooled and self.lazy_march and o.shape[0] <= self.n_rays:
            # lazy marching: only the head of every ray is independent of the model (the tail follows the visibility of the head)
            return ops.march_head(o, d, self._aabb_host, g._roi_host, g._resolution, g._binary_u8(), self.near, self.far, self.step_size,
                                  k0=min(32, self.early_termination), pool=self.pool_bufs)
        cap = ops.march_capacity(o.shape[0], self.near, self.far, self.step_size) if bf16 else None
        # sync-free mode: pooled arrays, two alternating sets because a prefetched batch is alive next to the current one
        pool, tag = None, "march"
        if pooled and self.sync_free and o.shape[0] <= self.n_rays:
            pool, tag = self.pool_bufs, "march%d" % (self._march_calls % 2)
            self._march_calls += 1
        return ops.march(o, d, self._aabb_host, g._roi_host, g._resolution, g._binary_u8(), self.near, self.far, self.step_size,
                         capacity=cap, total_out=totals[0:1] if bf16 else None, pool=pool, tag=tag)

    def _draw(self, march):
        """Next ray batch from the pool (+ its march when the occupancy grid will not change before it is used)."""
        totals = torch.zeros((4,), dtype=torch.int32, device=self.dev)     # [marched, kept, sampler candidates, sampler overflow]
        stream = self.pool._seed_streams.get(id(self.ray_gen))
        pre_state = stream.bit_generator.state if stream is not None else None   # a checkpoint taken while this batch is pending re-draws it
        o, d, target = self.pool.sample(self.n_rays, generator=self.ray_gen, status=totals[2:4])
        return dict(o=o, d=d, target=target, totals=totals, marched=self._march(o, d, totals) if march else None, pre_state=pre_state,
                    had_stream=stream is not None)

    def march_and_filter(self, o, d, totals=None, sync_free=None, marched=None):
        """acc_ray_marching (run_nerf_acc.py:287): returns (ray_idx int32, t0, t1, offsets, n_prefilter).

        fp32 check path: exact-size arrays, two host syncs (marched and kept counts), like the reference library.
        bf16 path: the marcher writes into capacity-sized arrays and leaves its sample count on the device, the
        visibility-pass MLP reads it there.  With sync_free the compaction does the same (no host sync at all; n_prefilter is
        returned as None and all counts stay in `totals` = [marched, kept, sampler candidates, sampler overflow]); otherwise
        ONE read-back after the visibility scan sizes the training buffers and returns every counter at once."""
        g = self.acc_grid
        R = o.shape[0]
        bf16 = self.model._precision_id == ops.PREC_BF16
        sync_free = self.sync_free if sync_free is None else (sync_free and bf16)
        cap = ops.march_capacity(R, self.near, self.far, self.step_size) if bf16 else None
        if bf16 and totals is None:
            totals = torch.zeros((4,), dtype=torch.int32, device=o.device)
        if sync_free and self.lazy_march:                # `marched`, if given, is the head march of these rays (see _march)
            ray_idx, t0, t1, offsets = ops.march_filter_lazy(
                self.model._desc, self.kflat, self.packed, self.model._precision_id, o, d, self._aabb_host, g._roi_host, g._resolution,
                g._binary_u8(), self.near, self.far, self.step_size, self.early_stop_eps, self.alpha_thre,
                k0=min(32, self.early_termination), totals=totals, pool=self.pool_bufs, timing=self.kernel_events, head=marched,
                thre_cap=g.occs_mean_dev())
            return ray_idx, t0, t1, offsets, None
        ray_idx, t0, t1, offsets = marched if marched is not None else self._march(o, d, totals, pooled=sync_free)
        n_pre = ray_idx.numel()
        if n_pre > 0:
            if self.early_termination:
                # same kept set as evaluating every sample; only samples that can still be visible reach the MLP
                alphas, _ = ops.alphas_two_phase(self.model._desc, self.kflat, self.packed, self.model._precision_id, o, d, ray_idx,
                                                 t0, t1, offsets, self.early_stop_eps, k0=self.early_termination, timing=self.kernel_events,
                                                 pool=self.pool_bufs if sync_free else None)
            else:
                if self.kernel_events is not None:
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record()
                alphas = ops.mlp_forward(self.model._desc, self.kflat, self.packed, ops.OUT_ALPHA, self.model._precision_id,
                                         rays_o=o, rays_d=d, ray_idx=ray_idx, t_starts=t0, t_ends=t1,
                                         n_dev=offsets[R:R + 1] if bf16 else None)
                if self.kernel_events is not None:
                    ev1.record()
                    self.kernel_events.append((ev0, ev1, totals[0:1] if bf16 else n_pre))
            # alpha_thre = min(alpha_thre, mean(occs)) as nerfacc.ray_marching does; the mean is read on the device
            ray_idx, t0, t1, offsets, host_totals = ops.visibility_compact(alphas, offsets, t0, t1, self.early_stop_eps, self.alpha_thre,
                                                                          thre_cap=g.occs_mean_dev(), totals=totals if bf16 else None,
                                                                          capacity=cap if sync_free else None,
                                                                          pool=self.pool_bufs if sync_free else None)
            if sync_free:
                n_pre = None
            elif bf16:
                n_pre = host_totals[0]
                if host_totals[3] != 0:
                    raise RuntimeError("ray sampler: candidate buffer overflow / underflow -- re-draw with a larger threshold")
        return ray_idx, t0, t1, offsets, n_pre

    def _refresh_packed(self):
        """Parameters as the kernels see them this step (the master buffer itself unless the model folds a BARF mask into its
        first layer) and, on the bf16 path, their packed bf16 image."""
        self.kflat = self.model._kernel_params()
        if self.model._precision_id == ops.PREC_BF16:
            self.packed = ops.mlp_pack(self.model._desc, self.kflat, self.packed)

    # ------------------------------------------------------------------ one reference iteration
    def step(self, rays=None):
        """rays: optional (o[R,3], d[R,3], target[R]) -- otherwise sampled from the pool.  Returns a StepResult: `loss` is a
        device scalar, the sample counts are fetched lazily.  In sync-free mode nothing in here waits for the GPU: mean(occs) of
        a grid refresh stays on the device (the visibility kernels read it there); only the refreshes after the grid's warm-up
        (iteration >= 256) pick their cells with torch.nonzero, which synchronises like the reference library does.

        Step boundary (sync-free mode, pool sampling): the NEXT iteration's ray batch is drawn and marched between this
        iteration's backward and its optimiser step -- neither depends on the weights -- so the gradient all-reduce runs
        underneath ~0.4 ms of independent work and a rank that finishes its backward early does not stall on the slowest
        rank.  The sequence of sampler draws, and therefore every result, is the same as without the reordering."""
        m = self.model
        if self.flat.data_ptr() != m._flat.data_ptr():
            raise RuntimeError("model parameters were re-allocated after the Trainer was built")
        marched = None
        if rays is None:
            batch, self._prefetched = self._prefetched, None
            if batch is None:
                batch = self._draw(march=False)
            o, d, target, totals, marched = batch["o"], batch["d"], batch["target"], batch["totals"], batch["marched"]
        else:
            o, d, target = rays
            totals = torch.zeros((4,), dtype=torch.int32, device=self.dev)
            if self._prefetched is not None and self.lazy_march:
                # the pending batch's head march lives in the pooled lz_h_* arrays this step is about to reuse: redo it later
                self._prefetched["marched"] = None
        R = o.shape[0]
        sync_free = self.sync_free and R <= self.n_rays
        if rays is None and sync_free:
            self._check_sampler_status(totals)
        self._refresh_packed()
        self.update_grids()
        ray_idx, t0, t1, offsets, n_pre = self.march_and_filter(o, d, totals, sync_free=sync_free, marched=marched)
        n_kept = ray_idx.numel()                                            # sync-free: the capacity, not the count
        have = n_kept > 0                                                   # run_nerf_acc.py:289 (rank-local on the one-sync path)
        # Data-parallel runs enter the gradient exchange on EVERY iteration: `have` is a rank-local fact on the one-sync path, so
        # a rank without samples contributes a zero gradient and a zero count instead of skipping the collective its peers
        # wait in; the all-rank sum of the count slot gates Adam on every rank alike.
        if have or self.world > 1:
            prec = m._precision_id
            use_peer = self.peer is not None and sync_free
            if use_peer:
                self.grad, tag = self.peer.next_buffer()                    # this step's gradient lives in NVLink peer memory
            if have:
                kw = dict(rays_o=o, rays_d=d, ray_idx=ray_idx, t_starts=t0, t_ends=t1)
                if sync_free:
                    kw["n_dev"] = offsets[R:R + 1]
                chunk = None if sync_free else self._backward_chunk(n_kept)
                if chunk is None:
                    logits, saved = ops.mlp_forward(m._desc, self.kflat, self.packed, ops.OUT_LOGIT, prec, saved=True, pool=self.pool_bufs, **kw)
                    pix, glogits, loss_sum = ops.composite_mse_fused(logits, t0, t1, offsets, target, R * self.world,
                                                                     pool=self.pool_bufs if sync_free else None)
                    ops.mlp_backward(m._desc, self.kflat, self.packed, saved, glogits, prec, grad_params=self.grad, pool=self.pool_bufs, **kw)
                else:
                    # The saved tile images of ALL kept samples would not fit (8 x 256: 9.4 KB per sample): logits of every sample
                    # first (no activations kept), composite + loss, then forward(train) -> dgrad -> wgrad chunk by chunk with
                    # the parameter gradient accumulated over the chunks.  Same numbers as the one-shot path up to the fp32
                    # summation order of the weight gradients.
                    logits = ops.mlp_forward(m._desc, self.kflat, self.packed, ops.OUT_LOGIT, prec, **kw)
                    pix, glogits, loss_sum = ops.composite_mse_fused(logits, t0, t1, offsets, target, R * self.world)
                    self.grad.zero_()
                    if self._grad_chunk is None:
                        self._grad_chunk = torch.empty_like(self.grad)
                    for a in range(0, n_kept, chunk):
                        b = min(n_kept, a + chunk)
                        kc = dict(rays_o=o, rays_d=d, ray_idx=ray_idx[a:b], t_starts=t0[a:b], t_ends=t1[a:b])
                        _, saved = ops.mlp_forward(m._desc, self.kflat, self.packed, ops.OUT_LOGIT, prec, saved=True, pool=self.pool_bufs, **kc)
                        ops.mlp_backward(m._desc, self.kflat, self.packed, saved, glogits[a:b], prec, grad_params=self._grad_chunk,
                                         pool=self.pool_bufs, **kc)
                        self.grad[:-1].add_(self._grad_chunk[:-1])
                m._map_grad_(self.grad)                                     # BARF: chain rule through the folded mask (no-op otherwise)
                loss = loss_sum / R
            else:
                self.grad.zero_()
                loss = torch.full((1,), float("nan"), device=self.dev)
                pix = torch.ones(R, device=self.dev)
            active = None
            if sync_free:
                self.grad[-1:].copy_(offsets[R:R + 1])                      # kept count rides behind the gradient
                active = self.grad[-1:]
            elif self.world > 1:
                self.grad[-1:].fill_(float(n_kept))
                active = self.grad[-1:]
            work = None
            if use_peer:
                ops.signal_peers(self.peer, tag)                            # gradient complete -> tag to every peer
            elif self.world > 1:                                            # sum of per-rank (1/global-batch)-scaled grads
                work = torch.distributed.all_reduce(self.grad, group=self.pg, async_op=True)
            if rays is None and sync_free and self.prefetch:
                # the occupancy grid is refreshed at the START of iterations that are multiples of 16: march ahead only otherwise
                self._prefetched = self._draw(march=(self.n_iter + 1) % self.GRID_EVERY != 0)
            if work is not None:
                work.wait()
            # n_iter_adam counts optimiser calls; an iteration in which NO rank kept a sample is skipped on the device (`active`)
            # but still counted here -- it would take 65 536 rays without a single kept sample, and reading the gate back would
            # cost the host synchronisation this loop exists to avoid.
            if use_peer:                                                    # waits for all tags, sums over NVLink, Adam -- one kernel
                ops.adam_step_allreduce(self.flat, self.peer, tag, self.exp_avg, self.exp_avg_sq, self.lr, self.n_iter_adam + 1,
                                        active_index=self.flat.numel(), wait_stats=self.peer.wait_stats)
            else:
                ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.n_iter_adam + 1, active=active)
            self.n_iter_adam += 1
            self.lr = self.lr0 * (self.decay_rate ** (self.n_iter / self.decay_steps))   # run_nerf_acc.py:323-328
        else:
            loss = torch.full((1,), float("nan"), device=self.dev)
            pix = torch.ones(R, device=self.dev)
        self.n_iter += 1
        self.last = StepResult(loss=loss, totals=totals, n_rays=R, pix=pix)
        if not sync_free:
            self.last["n_samples_prefilter"], self.last["n_samples"] = n_pre, n_kept
        return self.last

    n_iter_adam = 0
    _grad_chunk = None

    def _backward_chunk(self, n_kept):
        """Samples per forward(train)/backward chunk on the one-sync path, or None when the saved tile images + backward
        workspace of all n_kept samples fit in `train_memory_bytes` (default: 40 % of the HBM that was free at construction)."""
        if self._bytes_per_sample is None:
            lib, desc, prec = _lib.load(), self.model._desc, self.model._precision_id
            probe = 1 << 20
            self._bytes_per_sample = (int(lib.angio_mlp_saved_bytes(ctypes.byref(desc), probe, prec)) +
                                      int(lib.angio_mlp_workspace_bytes(ctypes.byref(desc), probe, prec, 1))) / probe
        if n_kept * self._bytes_per_sample * 2.1 <= self.train_memory_bytes:      # the buffer pool doubles when it grows
            return None
        chunk = int(self.train_memory_bytes / (2.1 * self._bytes_per_sample)) // 128 * 128
        return max(chunk, 128 * 148)

    _bytes_per_sample = None

    def _check_sampler_status(self, totals):
        """The on-device ray sampler reports a failed draw (candidate buffer overflow / under-fill; its ids are then all ray 0)
        in totals[3].  The sync-free loop never reads it, so every 16th step it is copied to pinned host memory without
        waiting, and a later step raises once that copy has completed."""
        ev = self._status_event
        if ev is not None and ev.query():
            self._status_event = None
            if int(self._status_host[1]) != 0:
                raise RuntimeError("ray sampler: candidate buffer overflow / underflow in an earlier draw -- re-draw with a larger threshold")
        if self._status_event is None and self._status_host is not None and self.n_iter % self.GRID_EVERY == 0:
            self._status_host.copy_(totals[2:4], non_blocking=True)
            self._status_event = torch.cuda.Event()
            self._status_event.record()

    # ------------------------------------------------------------------ state capture (benchmark arms start from the same state)
    def snapshot(self):
        grids = [g for g in (self.acc_grid, self.vessel_acc_grid) if g is not None]
        return dict(flat=self.flat.clone(), m=self.exp_avg.clone(), v=self.exp_avg_sq.clone(), n_iter=self.n_iter,
                    n_iter_adam=self.n_iter_adam, lr=self.lr, grids=[(g.occs.clone(), g._binary.clone(), g.occs_mean_host) for g in grids],
                    gens=(self.ray_gen.get_state(), self.grid_gen.get_state()), prefetched=self._prefetched,
                    seed_stream=(self.pool._seed_streams[id(self.ray_gen)].bit_generator.state
                                 if id(self.ray_gen) in self.pool._seed_streams else None))

    def restore(self, snap):
        self.flat.copy_(snap["flat"]); self.exp_avg.copy_(snap["m"]); self.exp_avg_sq.copy_(snap["v"])
        self.n_iter, self.n_iter_adam, self.lr = snap["n_iter"], snap["n_iter_adam"], snap["lr"]
        grids = [g for g in (self.acc_grid, self.vessel_acc_grid) if g is not None]
        for g, (occs, binary, mean) in zip(grids, snap["grids"]):
            g.occs.copy_(occs); g._binary.copy_(binary); g.occs_mean_host = mean
        self.ray_gen.set_state(snap["gens"][0]); self.grid_gen.set_state(snap["gens"][1])
        # the pending batch keeps its rays (own tensors); its march lived in pooled arrays that later steps have reused -> redo it
        pre = snap.get("prefetched")
        self._prefetched = dict(pre, marched=None) if pre is not None else None
        if snap.get("seed_stream") is not None:
            rng = np.random.default_rng(0)
            rng.bit_generator.state = snap["seed_stream"]
            self.pool._seed_streams[id(self.ray_gen)] = rng

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, filename, extra=None):
        """One .pth in the reference's checkpoint layout (CPPN.save, /root/reference/model/CPPN.py:261-276: version /
        parameters / training_information / model state_dict) whose `training_information` additionally carries what a true
        resume needs and the reference never stored: Adam moments, step counters, lr, both occupancy grids, RNG states.
        A checkpoint written here loads in the reference (it only reads 'parameters' and 'model'), and vice versa."""
        grids = [g for g in (self.acc_grid, self.vessel_acc_grid) if g is not None]
        info = dict(extra or {})
        info["resume"] = dict(
            n_iter=self.n_iter, n_iter_adam=self.n_iter_adam, lr=self.lr,
            exp_avg=self.exp_avg.detach().cpu(), exp_avg_sq=self.exp_avg_sq.detach().cpu(),
            grids=[dict(occs=g.occs.detach().cpu(), binary=g._binary.detach().cpu(), occs_mean=g.occs_mean_host) for g in grids],
            ray_gen=self.ray_gen.get_state().cpu(), grid_gen=self.grid_gen.get_state().cpu(),
            sampler_seed_stream=(self._prefetched["pre_state"] if self._prefetched is not None else
                                 (self.pool._seed_streams[id(self.ray_gen)].bit_generator.state
                                  if id(self.ray_gen) in self.pool._seed_streams else None)))
        self.model.save(filename, info)

    def save_grids(self, log_dir, prefix="coarse"):
        """`pv_grid.save(f"{log_dir}coarsegrid.vtk")` / `...coarsevesselgrid.vtk` (run_nerf_acc.py:359-367; prefix 'high' at
        :384-385): the occupancy grids in the reference's legacy-VTK layout, readable by its inference script."""
        from . import gridio
        gridio.save_grid_vtk(os.path.join(log_dir, f"{prefix}grid.vtk"), self.acc_grid)
        if self.vessel_acc_grid is not None:
            gridio.save_grid_vtk(os.path.join(log_dir, f"{prefix}vesselgrid.vtk"), self.vessel_acc_grid)

    def load_checkpoint(self, filename):
        """Restore model + optimiser + grids + counters from save_checkpoint (or just the model from a reference .pth)."""
        ck = torch.load(filename, map_location="cpu", weights_only=False)
        self.model.load_state_dict(ck["model"])
        if self.flat.data_ptr() != self.model._flat.data_ptr():
            raise RuntimeError("load_state_dict re-allocated the flat parameter buffer")
        res = (ck.get("training_information") or {}).get("resume")
        if res is None:
            return ck
        self.n_iter, self.n_iter_adam, self.lr = res["n_iter"], res["n_iter_adam"], res["lr"]
        self.exp_avg.copy_(res["exp_avg"]); self.exp_avg_sq.copy_(res["exp_avg_sq"])
        grids = [g for g in (self.acc_grid, self.vessel_acc_grid) if g is not None]
        for g, st in zip(grids, res["grids"]):
            g.occs.copy_(st["occs"]); g._binary.copy_(st["binary"]); g.occs_mean_host = st["occs_mean"]
        self.ray_gen.set_state(res["ray_gen"]); self.grid_gen.set_state(res["grid_gen"])
        self._prefetched = None
        if res.get("sampler_seed_stream") is not None:               # host-side seed stream of the on-device ray sampler
            rng = np.random.default_rng(0)
            rng.bit_generator.state = res["sampler_seed_stream"]
            self.pool._seed_streams[id(self.ray_gen)] = rng
        return ck

    # ------------------------------------------------------------------ evaluation (run_nerf_acc.py:338-357)
    @torch.no_grad()
    def render_view(self, v, binary_thresh=None):
        o, d, target = self.pool.rays_of_view(v)
        self._refresh_packed()
        ray_idx, t0, t1, offsets, _ = self.march_and_filter(o, d, sync_free=False)    # exact-size arrays for the eval path
        if ray_idx.numel() == 0:
            return torch.ones_like(target), target
        logits = ops.mlp_forward(self.model._desc, self.kflat, self.packed, ops.OUT_LOGIT, self.model._precision_id,
                                 rays_o=o, rays_d=d, ray_idx=ray_idx, t_starts=t0, t_ends=t1)
        zero = None
        if binary_thresh is not None:
            zero = (torch.sigmoid(logits) < binary_thresh).to(torch.uint8)
        return ops.composite_forward(logits, t0, t1, offsets, zero), target

    @torch.no_grad()
    def evaluate(self, v=None):
        """PSNR of the test view (the LAST view of the pool, run_nerf_acc.py:85,355-357) and, when the pool carries a
        sampling-weight image, the vessel-pixel PSNR over the test-view pixels whose weight exceeds the view's mean weight
        (run_nerf_acc.py:102-105,370-374; None when there is no such pixel or no weight image)."""
        v = self.pool.n_views - 1 if v is None else v
        pix, target = self.render_view(v)
        mse = torch.mean((pix - target) ** 2)
        vessel_psnr = None
        w = getattr(self.pool, "weights", None)
        if w is not None:
            wv = w[int(v)].reshape(-1)
            sel = wv > wv.mean()
            if bool(sel.any()):
                vessel_psnr = float(-10.0 * torch.log10(torch.mean((pix[sel] - target[sel]) ** 2)))
        return dict(mse=float(mse), psnr=float(-10.0 * torch.log10(mse)), vessel_psnr=vessel_psnr,
                    image=pix.view(self.pool.img_h, self.pool.img_w))


def psnr_from_loss(loss: float) -> float:
    return -10.0 * math.log10(loss)
