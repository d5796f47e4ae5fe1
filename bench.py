#!/usr/bin/env python
"""bench.py -- throughput benchmark of the NeRF-for-angiography hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload config3|config2|config4|config5|tiny] [--advance A] [--repeats P]

Training workloads (config2/3/4, tiny).  One "step" = one reference training iteration
(/root/reference/nerf/run_nerf_acc.py:263-328): draw a ray batch with the reference's distance-weighted sampling, refresh the
occupancy grids (every 16th step), march + visibility filter, MLP forward, Beer-Lambert composite + MSE, backward, Adam.
Default workload: BASELINE.json configs[2] ("config3": CT-derived phantom, 512x512 cone-beam, 60 views + test view, 4x128 Fourier
MLP, occupancy-grid marching, 65 536 rays per GPU per step), random-init weights.

What is timed.  The workload changes while the network trains: at random init the field is opaque and every ray dies inside its
first 32 samples (cheap steps); after a few dozen iterations the field has thinned out, nothing terminates early and the step is
~4x more expensive -- and stays there.  `value` is therefore measured in that STEADY regime: after the W warm-up steps the run
advances A (default 300) untimed real iterations, takes a snapshot, and times the K-step window P (default 5) times from the same
snapshot; `value` / `ms_per_step` are the median window, `spread` its min / max.  The cheap early window (iterations W..W+K) is
reported beside it as `early_window`.

Keys of the ONE JSON line (rank 0): `value` = rays/s with the step's inputs resident in HBM; `e2e` = the same metric with each
step's ray batch arriving from pinned HOST memory and the loss read back (the reference samples rays on the host and copies them
every iteration, nerf/nerf_helpers.py:144-148); `roofline` = the dominant kernel (fused tcgen05 MLP forward of the no-grad
visibility pass) against the measured bf16 peaks; `rooflines` = every kernel of the step with its CUDA-event time per step, its
algorithmic bytes / FLOPs and the fraction of the measured HBM / bf16 peaks; `cpu_baseline` = the oracle port timed on this box's
host cores; `allreduce_wait_us` = how long the fused all-reduce + Adam kernel waited for its slowest peer (N > 1).

Inference workload (config5, BASELINE.json configs[4]): one step = every GPU renders `views_per_step` 512x512 novel views through
the same hot path (visualization/visualization.py:335-352) after A training iterations; a 512^3 attenuation-volume query
(visualization.py:209-229) sharded by slabs is timed beside it.
"""
import argparse
import ctypes
import gc
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: detector size, theta sweep (+ the test view), rays/GPU/step, phantom kind / resolution, MLP
    "config3": dict(img=512, thetas=[6.0 * i for i in range(60)], rays=65536, vol=256, kind="ct_hu", L=4, H=128, enc="fourier"),
    "config2": dict(img=256, thetas=[0.0, 45.0, 90.0, 135.0], rays=65536, vol=256, kind="ct", L=4, H=128, enc="fourier"),
    # BASELINE configs[3]: 1024^2 x 120 views, 8x256 MLP
    "config4": dict(img=1024, thetas=[3.0 * i for i in range(120)], rays=131072, vol=256, kind="ct_hu", L=8, H=256, enc="fourier"),
    # BASELINE configs[4]: inference -- 360 novel views at 512^2 + a 512^3 volume query, trained 4x128 model + grid
    "config5": dict(img=512, thetas=[6.0 * i for i in range(60)], rays=65536, vol=256, kind="ct_hu", L=4, H=128, enc="fourier",
                    inference=dict(views=360, views_per_step=4, volume=512)),
    "tiny": dict(img=64, thetas=[22.5 * i for i in range(8)], rays=4096, vol=64, kind="ct", L=4, H=128, enc="fourier"),
}
MLP_FWD_FLOP = {("fourier", 4, 128): 139776, ("none", 4, 128): 132096, ("fourier", 8, 256): 1065984}
MLP_PARAMS = {("fourier", 4, 128): 70544, ("none", 4, 128): 66689, ("fourier", 8, 256): 535312}


def model_def(w, device, precision):
    return {'num_early_layers': w["L"], 'num_late_layers': 0, 'num_filters': w["H"], 'num_input_channels': 3,
            'num_output_channels': 1, 'num_input_channels_views': 0, 'use_bias': True, 'pos_enc': w["enc"], 'pos_enc_basis': 5,
            'act_func': 'relu', 'fourier_sigma': 5, 'num_img': 1, 'device': device, 'precision': precision}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (nvidia_ml_py) -- the counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints.  The samples are taken inline by the timing loop
    (every few steps, while the GPU is busy with the steps already enqueued): a background Python thread would contend for the
    GIL with the thread that launches kernels (5 ms switch interval = visible stalls in a 2.4 ms step), and a polling
    nvidia-smi process perturbs the driver.  NVML is loaded and exercised once before the timed region."""

    def __init__(self, index):
        self.rows, self.index, self.nvml, self.handle, self.sm_max, self.bits = [], index, None, None, None, {}

    def prepare(self):
        try:
            import pynvml as n
            n.nvmlInit()
            self.handle = n.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            self.bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                         "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
            self.nvml = n
            self.sample()
            self.rows.clear()
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is None:
            return
        try:
            sm = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            r = int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.rows.append((sm, [k for k, b in self.bits.items() if r & b]))
        except Exception:
            pass

    def result(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted({k for r in self.rows for k in r[1]}),
                "samples": len(sm), "source": "nvml, sampled inline while the timed steps execute"}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_oracle_step_time(w, n_rays_cpu, steps=1, seed=0):
    """Time the ORACLE (CPU restatement of the reference path, torch-CPU fp32 + C marcher) on a bounded sample of the
    same workload: `n_rays_cpu` rays of view 0 against a fully occupied grid (the state the GPU benchmark is in)."""
    import functools
    from oracle import cppn as ocppn, geometry as ogeo, nerfacc_ref, pipeline
    torch.set_num_threads(os.cpu_count() or 1)
    p = ocppn.init_params(w["L"], w["H"], w["enc"], 5, 5.0, seed=seed)
    params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f = functools.partial(ocppn.cppn_forward, params, pos_enc=w["enc"], basis=5)
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
    grid = nerfacc_ref.OccupancyGrid(roi, 128)
    grid.binary[:] = True
    grid.occs[:] = 0.5
    W = w["img"]
    o, d, _ = ogeo.get_ray_values(0.0, 0.0, 0.0, [0, 0, 1500.0], W, W, 7.5 * W)
    rng = np.random.default_rng(seed)
    sel = rng.permutation(W * W)[:n_rays_cpu]
    o = o.reshape(-1, 3)[sel].astype(np.float32)
    d = d.reshape(-1, 3)[sel].astype(np.float32)
    target = torch.from_numpy(rng.random(n_rays_cpu).astype(np.float32))
    times, n_pre, n_kept = [], 0, 0
    for _ in range(steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            ri, ts, te, pre = pipeline.acc_ray_marching(f, grid, roi, o, d, 300, 1400.0, 1600.0, 1e-2, 1e-4, return_prefilter=True)
        pos = pipeline.midpoints(torch.from_numpy(o), torch.from_numpy(d), ri, torch.from_numpy(ts), torch.from_numpy(te))
        pred = pipeline.get_predictions(f, pos, 131072)
        pix = pipeline.acc_render_volume_density(pred, ri, torch.from_numpy(ts), torch.from_numpy(te), n_rays_cpu)
        loss = torch.nn.functional.mse_loss(pix, target)
        opt.zero_grad()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
        n_pre, n_kept = len(pre[0]), len(ri)
    return float(np.mean(times)), n_pre, n_kept


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; nerfacc is CUDA-only and absent,
    see DESIGN.md) on the host cores, bounded sample per step."""
    if rank != 0:
        return
    big = args.workload != "tiny"
    n_cpu = (2048 if w["H"] > 128 else 8192) if big else 256          # ~1.5 s of CPU work per step on 16 cores
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_step_time(w, n_cpu, 1)
    n_timed = max(1, min(args.steps, 8)) if big else args.steps       # bounded: the whole arm ends within a few minutes
    t, n_pre, n_kept = cpu_oracle_step_time(w, n_cpu, n_timed)
    val = n_cpu / t
    line = {"impl": "reference", "metric": "train_rays_per_s", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": n_timed,
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sample": f"{n_cpu} rays/step of view 0, full 128^3 grid, {n_pre} marched / {n_kept} kept samples, "
                                                             f"{n_timed} steps timed"},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n_timed} training steps of {n_cpu} rays (oracle port: torch-CPU fp32 MLP + C marcher)"},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ per-kernel rooflines
def kernel_timeline(lib, run_steps, n_steps):
    """Device time of every kernel launch of the library over `run_steps()` (CUDA events recorded at each launch, see
    angio_profile_start): {kernel name: (ms per step, launches per step)}."""
    lib.angio_profile_start(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    run_steps()
    n = int(lib.angio_profile_stop())
    if n < 0:
        raise RuntimeError("angio_profile_stop failed")
    agg = {}
    name = ctypes.create_string_buffer(96)
    ms = ctypes.c_float()
    for i in range(n):
        if lib.angio_profile_entry(i, name, 96, ctypes.byref(ms)) != 0:
            raise RuntimeError("angio_profile_entry failed")
        k = name.value.decode()
        t, c = agg.get(k, (0.0, 0))
        agg[k] = (t + float(ms.value), c + 1)
    return {k: (t / n_steps, c / n_steps) for k, (t, c) in agg.items()}


def build_rooflines(timeline, cnt, w, peaks, world):
    """cnt: per-step averages {rays, head, tail, kept, pool, params, grid_cells}.  Algorithmic work per unit: SURVEY.md 8(d)."""
    flop = MLP_FWD_FLOP.get((w["enc"], w["L"], w["H"]), 0)
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    burst = float(peaks.get("bf16_tflops", 1590.0))
    sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    R, head, tail, kept = cnt["rays"], cnt["head"], cnt["tail"], cnt["kept"]
    marched = head + tail
    H = w["H"]
    img_bytes = ((3 + 6 * 5 + 15) // 16 * 16 + (w["L"] + 1) * H) * 2            # a_0 .. a_{L+1} bf16 tile images per sample
    delta_bytes = (w["L"] + 1) * H * 2
    spec = {   # kernel: (bound, algorithmic work per step, unit, note)
        "mlp_fwd_tc_kernel<ALPHA>": ("tensor", flop * cnt.get("evals", marched), "no-grad visibility pass: 2*MAC per evaluated sample"),
        "mlp_fwd_tc_kernel<LOGIT,train>": ("tensor", flop * kept, f"training forward, 2*MAC per kept sample; also streams {img_bytes + 16 * (w['L'] + 1)} B/sample of saved tile images"),
        "mlp_dgrad_tc_kernel": ("tensor", flop * kept, f"data-gradient chain, 2*MAC per kept sample; also streams {delta_bytes} B/sample of delta images"),
        "mlp_wgrad_tc_kernel": ("tensor", flop * kept, f"weight gradients, 2*MAC per kept sample; streams {img_bytes + delta_bytes} B/sample (HBM-bound by design)"),
        "mlp_fwd_tc_kernel<SIGMA>": ("tensor", flop * cnt["grid_cells"], "occupancy-grid refresh (amortised over 16 steps)"),
        "march_head_kernel": ("hbm", 24 * R + 12 * head, "24 B/ray in, 12 B/head sample out"),
        "march_count_warp_kernel": ("hbm", 28 * R + 72 * R, "warp per ray, 32 candidate samples per grid look-up round (latency-bound): 24 B/ray in, 4 B/ray + run table out"),
        "march_write_runs_kernel": ("hbm", 12 * tail + 4 * R, "warp per ray from the run table: 12 B/tail sample out"),
        "march_count_kernel": ("hbm", 8 * R, "stores t_min / t_max; serial walk only for rays whose t-chain crosses an fp32 binade (none here)"),
        "visibility_head_mask_kernel": ("hbm", 5 * head + 12 * R, "4 B/sample in, 1 B/sample out"),
        "visibility_mask_kernel": ("hbm", 5 * tail + 8 * R, "4 B/sample in, 1 B/sample out"),
        "compact_head_tail_kernel": ("hbm", 1 * marched + 20 * kept + 8 * R, "1 B/marched sample + 8 B/kept in, 12 B/kept out"),
        "composite_mse_kernel": ("hbm", 16 * kept + 12 * R, "12 B/kept sample in, 4 B/kept sample out"),
        "outgrad_partial_kernel": ("hbm", (2 * H + 4) * kept, "output-layer gradient: re-reads the a_{L+1} tile images"),
        "sample_candidates_kernel": ("hbm", 4 * cnt["pool"], "4 B/pool ray (weight image)"),
        "raygen_flat_kernel": ("hbm", 40 * R, "8 B id + 4 B pixel in, 28 B/ray out"),
        "adam_kernel": ("hbm", 28 * cnt["params"], "16 B/param in, 12 B/param out"),
        "adam_allreduce_kernel": ("nvlink-latency", (4 * world + 24) * cnt["params"], "world x 4 B/param peer reads + Adam"),
    }
    for a, b in (("mlp_fwd_tc_kernel", "mlp256_fwd_kernel"), ("mlp_dgrad_tc_kernel", "mlp256_dgrad_kernel"), ("mlp_wgrad_tc_kernel", "mlp256_wgrad_kernel"),
                 ("outgrad_partial_kernel", "outgrad256_partial_kernel")):          # the width-256 kernel family: same accounting
        for k in [k for k in spec if k.startswith(a)]:
            spec[k.replace(a, b)] = spec[k]
    if tail == 0.0:       # one-sync path: full march (count + write) and a two-phase visibility pass over it
        spec["march_write_runs_kernel"] = ("hbm", 12 * marched + 4 * R, "warp per ray from the run table: 12 B/sample out")
        spec["visibility_mask_kernel"] = ("hbm", 5 * marched + 8 * R, "4 B/sample in, 1 B/sample out")
        spec["compact_kernel"] = ("hbm", 1 * marched + 20 * kept + 8 * R, "1 B/marched sample + 8 B/kept in, 12 B/kept out")
    out = []
    for k, (ms, launches) in sorted(timeline.items(), key=lambda kv: -kv[1][0]):
        e = {"kernel": k, "ms_per_step": ms, "launches_per_step": launches}
        if k in spec and ms > 0:
            bound, work, note = spec[k]
            e["bound"], e["note"] = bound, note
            if bound == "tensor":
                ach = work / (ms * 1e-3) * 1e-12
                e.update(algorithmic_flop_per_step=work, achieved=ach, unit="TFLOP/s", frac_of_burst_peak=ach / burst, frac_of_sustained_peak=ach / sust)
            else:
                ach = work / (ms * 1e-3) * 1e-9
                e.update(algorithmic_bytes_per_step=work, achieved=ach, unit="GB/s", frac_of_hbm_peak=ach / hbm)
        else:
            e["bound"] = "latency"
        out.append(e)
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--advance", type=int, default=None, help="untimed real iterations between the warm-up and the timed (steady-regime) windows; default 300")
    ap.add_argument("--repeats", type=int, default=5, help="how many times the K-step window is timed from the same snapshot")
    ap.add_argument("--weights", default="distance", choices=["distance", "random"], help="ray-draw weights (reference: distance_pixel_value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample clocks during the timed region")
    ap.add_argument("--full-visibility", action="store_true",
                    help="evaluate the visibility-pass MLP on EVERY marched sample (the reference's order of operations) instead of the "
                         "two-phase pass with early ray termination; the kept samples are bit-identical either way")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if "precision" in w:
        args.precision = w["precision"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.advance is None:
        args.advance = 300 if args.workload != "tiny" else 40
    args.repeats = max(1, args.repeats)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import nerf_for_angiography_b200 as A
    from nerf_for_angiography_b200.data import make_dataset
    from nerf_for_angiography_b200.train import Trainer
    lib = A._lib.load()

    torch.manual_seed(0)
    pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], test_view=(135.0, 135.0), kind=w["kind"], volume_res=w["vol"],
                              device=dev, seed=0, weight_strategy=args.weights)
    model = A.CPPN(model_def(w, dev, args.precision)).to(dev)          # same seed => identical weights on every rank
    tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"], seed=0, early_termination=0 if args.full_visibility else 32)
    R = w["rays"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def timed_window(n_steps, body):
        """barrier + synchronize, n_steps x body(i) between two CUDA events, synchronize + barrier; GC off inside."""
        sync_all()
        gc.collect()
        gc.disable()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n_steps):
            body(i)
        b.record()
        sync_all()
        gc.enable()
        return a.elapsed_time(b)

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return t.tolist()

    if "inference" in w:
        run_inference(args, w, A, tr, info, dev, rank, world, sync_all, reduce_max)
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---------------- e2e batches: pinned host memory, one per step (reference: host-side sampling + H2D every iteration)
    host_batches = []
    for _ in range(args.steps + args.warmup):
        o, d, t = pool.sample(R, generator=tr.ray_gen)
        host_batches.append(tuple(x.cpu().pin_memory() for x in (o, d, t)))
    torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if not args.no_clocks:
        clocks.prepare()

    # ---------------- early window (iterations W .. W+K of a random-init network: opaque field, cheap steps)
    for _ in range(args.warmup):
        tr.step()
    sync_all()
    snap0 = tr.snapshot()
    for _ in range(args.steps):            # rehearsal: the caching allocator now owns every block the window needs
        tr.step()
    sync_all()
    tr.restore(snap0)
    early_totals = []
    ms_early = timed_window(args.steps, lambda i: early_totals.append(tr.step()["totals"]))
    ms_early = reduce_max([ms_early])[0]
    early_host = [t.tolist() for t in early_totals]

    # ---------------- advance into the steady regime, then time the window `repeats` times from one snapshot
    tr.restore(snap0)
    for _ in range(args.advance):
        tr.step()
    sync_all()
    occupied = float(tr.acc_grid.binary.float().mean())
    snap = tr.snapshot()
    for _ in range(args.steps):            # rehearsal (see above; also brings the grid refresh inside the window into the allocator)
        tr.step()
    sync_all()
    every = max(1, args.steps // 8)
    window_ms, step_totals, launches = [], [], 0
    use_peer = tr.peer is not None and tr.sync_free     # the one-sync loop (config 4) exchanges gradients with NCCL all_reduce
    if use_peer:
        tr.peer.wait_stats.zero_()
    for rep in range(args.repeats):
        tr.restore(snap)
        last = rep == args.repeats - 1
        if last:
            tr.kernel_events = []                                      # (start, end, sample count) per visibility-pass MLP launch
            launches0 = int(lib.angio_launch_count())
            torch.cuda.profiler.start()                                # ncu --profile-from-start off captures exactly one timed window

        def body(i, last=last):
            out = tr.step()
            if last:
                step_totals.append(out["totals"])
                if not args.no_clocks and i % every == every - 1:
                    clocks.sample()                                    # the GPU is executing the steps enqueued so far
            body.out = out
        window_ms.append(timed_window(args.steps, body))
        if last:
            torch.cuda.profiler.stop()
            launches = int(lib.angio_launch_count()) - launches0
    window_ms = reduce_max(window_ms)
    ms = float(np.median(window_ms))
    clk = clocks.result() if not args.no_clocks else None
    host_totals = [t.tolist() if isinstance(t, torch.Tensor) else list(t) for t in step_totals]
    if any(len(t) > 3 and t[3] != 0 for t in host_totals + early_host):
        raise RuntimeError("ray sampler overflow during the timed region")
    kernel_ms = [a.elapsed_time(b) for a, b, _ in tr.kernel_events]
    kernel_n = [int(c.item()) if isinstance(c, torch.Tensor) else int(c) for _, _, c in tr.kernel_events]   # samples each launch evaluated
    tr.kernel_events = None
    last_loss = float(body.out["loss"])
    wait = None
    if use_peer:
        ws = tr.peer.wait_stats.tolist()
        wt = torch.tensor([ws[0] / max(ws[1], 1) * 1e-3, ws[2] * 1e-3], dtype=torch.float64, device=dev)     # mean / longest wait of this rank, us
        mx, mean = wt.clone(), wt.clone()
        torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(mean, op=torch.distributed.ReduceOp.SUM)
        wait = {"mean_over_steps_us": {"max_over_ranks": float(mx[0]), "mean_over_ranks": float(mean[0]) / world},
                "longest_single_wait_us": float(mx[1]), "what": "time thread 0 of the fused all-reduce + Adam kernel spends waiting for the "
                "peers' step tags (globaltimer), i.e. for the slowest rank's backward"}

    # ---------------- per-kernel timeline of the same window (CUDA events at every launch; a separate, untimed pass)
    tr.restore(snap)
    kernel_timeline(lib, lambda: tr.step(), 1)          # untimed: creates the library's event pool outside the measured pass
    tr.restore(snap)
    sync_all()
    tl_totals, tr.kernel_events = [], []
    gc.collect()
    gc.disable()                                        # a collector pause would drain the stream and be billed to one kernel
    timeline = kernel_timeline(lib, lambda: [tl_totals.append(tr.step()["totals"]) for _ in range(args.steps)], args.steps)
    gc.enable()
    tl_host = [t.tolist() for t in tl_totals]
    tl_n = [int(c.item()) if isinstance(c, torch.Tensor) else int(c) for _, _, c in tr.kernel_events]
    tr.kernel_events = None
    kept_step = float(np.mean([t[1] for t in tl_host]))
    marched_step = float(np.mean([t[0] for t in tl_host]))
    if tr.lazy_march and len(tl_n) == 2 * args.steps:
        head_step, tail_step = float(np.mean(tl_n[0::2])), float(np.mean(tl_n[1::2]))
    else:
        head_step, tail_step = marched_step, 0.0
    n_refresh = sum(1 for i in range(snap["n_iter"], snap["n_iter"] + args.steps) if i % tr.GRID_EVERY == 0)
    cells = tr.acc_grid.num_cells * (2 if tr.vessel_acc_grid is not None else 1) * (1.0 if snap["n_iter"] < 256 else 0.5)
    counts = dict(rays=R, head=head_step, tail=tail_step, kept=kept_step, evals=float(sum(tl_n)) / args.steps, pool=pool.n_train_rays if pool.weights is not None else 0,
                  params=MLP_PARAMS.get((w["enc"], w["L"], w["H"]), 0), grid_cells=cells * n_refresh / args.steps)

    # ---------------- e2e arm: host buffers in, loss out, every step (same snapshot, median of 3 windows)
    tr.restore(snap)
    for i in range(args.warmup):
        o, d, t = (x.to(dev, non_blocking=True) for x in host_batches[i])
        float(tr.step(rays=(o, d, t))["loss"])
    e2e_ms = []
    for rep in range(min(3, args.repeats)):
        tr.restore(snap)

        def body_e2e(i):
            o, d, t = (x.to(dev, non_blocking=True) for x in host_batches[args.warmup + i])
            body_e2e.loss = float(tr.step(rays=(o, d, t))["loss"])    # D2H read of the step's loss
        e2e_ms.append(timed_window(args.steps, body_e2e))
    e2e_ms = reduce_max(e2e_ms)
    ms_e2e = float(np.median(e2e_ms))
    loss_host = body_e2e.loss

    cnt = torch.tensor([sum(t[0] for t in host_totals), sum(t[1] for t in host_totals), sum(t[0] for t in early_host),
                        sum(t[1] for t in early_host)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        flop = MLP_FWD_FLOP.get((w["enc"], w["L"], w["H"]))
        roofline = None
        if kernel_ms and flop and args.precision == "bf16":
            # every entry is one launch of the visibility-pass MLP forward (two per step with lazy marching: the first 32 samples
            # of every ray, then the rest of the rays still alive).  Launches that do not fill the 148 SMs twice are left out;
            # the rest are weighted by the samples they evaluated.
            big = [(n, t) for n, t in zip(kernel_n, kernel_ms) if t > 0 and n >= 2 * 128 * 148]
            if big:
                ach = flop * sum(n for n, _ in big) / (sum(t for _, t in big) * 1e-3) * 1e-12
                burst = float(peaks.get("bf16_tflops", 1590.0))
                sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
                roofline = {"bound": "tensor", "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst, "traffic": None,
                            "kernel": ("mlp256_fwd_kernel<ALPHA>" if w["H"] == 256 else "mlp_fwd_tc_kernel<ALPHA> [three-slot mlp_fwd3_tc_kernel]") +
                                      " (no-grad visibility pass, steady regime)",
                            "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; the stricter of the two measured peaks)" if peaks
                                            else "fallback 1.59 PFLOP/s"),
                            "frac_of_sustained_peak": ach / sust,
                            "avg_launch_ms": float(np.mean([t for _, t in big])), "samples_per_launch": float(np.mean([n for n, _ in big])),
                            "launches_timed": len(big), "flop_per_sample": flop}
                try:   # DRAM bytes of this kernel from the committed `ncu --set full` capture, scaled to this run's samples per launch
                    tr_ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["mlp_fwd_tc_kernel<ALPHA>"]
                    roofline["traffic"] = tr_ncu["dram_bytes"] / tr_ncu["samples"] * roofline["samples_per_launch"]
                    roofline["traffic_source"] = tr_ncu["source"]
                except Exception:
                    pass
        rooflines = build_rooflines(timeline, counts, w, peaks, world) if args.precision == "bf16" else None
        kernel_sum = sum(v[0] for v in timeline.values())
        cpu = None
        if not args.no_cpu_baseline:
            n_cpu, n_cpu_steps = ((2048 if w["H"] > 128 else 8192), 6) if args.workload != "tiny" else (256, 2)   # ~10 s of CPU work
            cpu_oracle_step_time(w, n_cpu, 1)                                          # untimed warm-up (thread pools, allocator)
            t_cpu, cp, ck = cpu_oracle_step_time(w, n_cpu, n_cpu_steps)
            cpu = {"value": n_cpu / t_cpu, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{n_cpu_steps} training steps of {n_cpu} rays of the oracle port ({cp} marched / {ck} kept samples per step, "
                             f"{t_cpu:.2f} s per step)"}
        rays = R * world * args.steps
        it0 = snap["n_iter"]
        line = {"metric": "train_rays_per_s", "value": rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "detector": f"{w['img']}x{w['img']}", "views": len(w["thetas"]) + 1,
                           "phantom": w["kind"], "ray_weights": args.weights,
                           "mlp": f"{w['L']}x{w['H']} {w['enc']}", "rays_per_gpu_per_step": R, "march_steps": 300, "grid": "128^3",
                           "iterations": [it0, it0 + args.steps],
                           "regime": f"steady: {args.advance} untimed training iterations after the warm-up; occupancy grid {occupied:.3f} occupied",
                           "windows": f"{args.repeats} x {args.steps} steps from one snapshot, median reported",
                           "l2": "no flush: every step draws fresh rays and streams the saved bf16 tile images of its kept samples "
                                 f"(~1 KB/sample x {float(cnt[1]) / world / max(args.steps, 1) / 1e6:.1f} M samples/step here) plus the sample arrays "
                                 "through HBM, far more than the 126 MB L2",
                           "visibility_pass": "every marched sample (reference order)" if args.full_visibility else
                                              "early ray termination: first 32 samples of every ray, then only the rays still transparent "
                                              + ("(marched lazily as well) " if tr.lazy_march else "") +
                                              "-- kept samples bit-identical to evaluating every sample"},
                "spread": {"window_ms": window_ms, "min_ms_per_step": min(window_ms) / args.steps, "max_ms_per_step": max(window_ms) / args.steps},
                "early_window": {"iterations": [args.warmup, args.warmup + args.steps], "ms_per_step": ms_early / args.steps,
                                 "value": rays / (ms_early * 1e-3), "samples_marched_per_step": float(cnt[2]) / world / args.steps,
                                 "samples_kept_per_step": float(cnt[3]) / world / args.steps,
                                 "what": "random-init network: opaque field, every ray terminates inside its first 32 samples"},
                "mlp_evals_visibility_per_step": float(sum(kernel_n)) / max(args.steps, 1),
                "visibility_launches": ({"head_samples_per_launch": float(np.mean(kernel_n[0::2])), "tail_samples_per_launch": float(np.mean(kernel_n[1::2])),
                                         "head_ms": float(np.mean(kernel_ms[0::2])), "tail_ms": float(np.mean(kernel_ms[1::2]))}
                                        if tr.lazy_march and len(kernel_n) == 2 * args.steps else None),
                "samples_marched_per_step": float(cnt[0]) / world / args.steps, "samples_kept_per_step": float(cnt[1]) / world / args.steps,
                "samples_per_s_marched": float(cnt[0]) / (ms * 1e-3), "samples_per_s_kept": float(cnt[1]) / (ms * 1e-3),
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": R * 28, "d2h_bytes_per_step": 4 + 8,
                        "window_ms": e2e_ms},
                "roofline": roofline, "rooflines": rooflines,
                "kernel_time_share": {"kernel_ms_per_step": kernel_sum, "step_ms": ms / args.steps, "share": kernel_sum / (ms / args.steps),
                                      "what": "sum of the per-launch CUDA-event times of one profiled window / the timed step"},
                "allreduce_wait_us": wait, "cpu_baseline": cpu, "final_loss": last_loss, "e2e_final_loss": loss_host}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def run_inference(args, w, A, tr, info, dev, rank, world, sync_all, reduce_max):
    """config 5: novel-view rendering + volume query with the model / grid after `--advance` training iterations."""
    from nerf_for_angiography_b200 import inference
    inf = w["inference"]
    for _ in range(args.warmup + args.advance):
        tr.step()
    sync_all()
    occupied = float(tr.acc_grid.binary.float().mean())
    model, grid = tr.model, tr.acc_grid
    img, vps = w["img"], inf["views_per_step"]
    all_views = [(360.0 * i / inf["views"], 0.0) for i in range(inf["views"])]
    src = np.array([0.0, 0.0, info["src_dist"]])
    kw = dict(src_pt=src, img_width=img, img_height=img, focal_length=7.5 * img, depth_samples_per_ray=300, near_thresh=info["near"],
              far_thresh=info["far"], early_stop_eps=1e-2, alpha_thre=1e-4, gather=False)
    my_views = all_views[rank::world]                      # interleaved: every rank sees the whole arc
    lib = A._lib.load()

    def render_step(i):
        vs = [my_views[(i * vps + k) % len(my_views)] for k in range(vps)]
        return inference.render_projections(model, grid, tr.scene_aabb, views=vs, shard=False, **kw)   # already this rank's share

    for i in range(max(3, args.warmup)):
        render_step(i)
    sync_all()
    launches0 = int(lib.angio_launch_count())
    window_ms = []
    for rep in range(args.repeats):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.steps):
            imgs = render_step(i)
        b.record()
        sync_all()
        window_ms.append(a.elapsed_time(b))
    launches = (int(lib.angio_launch_count()) - launches0) // args.repeats
    window_ms = reduce_max(window_ms)
    ms = float(np.median(window_ms))
    # e2e: the rendered images are copied to pinned host memory every step (the reference writes them to disk)
    host_img = torch.empty((vps, img, img), dtype=torch.float32).pin_memory()
    sync_all()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.steps):
        host_img.copy_(render_step(i), non_blocking=False)
    b.record()
    sync_all()
    ms_e2e = reduce_max([a.elapsed_time(b)])[0]
    # volume query: 512^3 lattice, slabs of the first axis sharded over the ranks
    n = inf["volume"]
    t = torch.linspace(-100.0, 100.0, n)
    inference.query_volume(model, t, grid=grid, gather=False)       # untimed: the caching allocator now owns the chunk buffers
    sync_all()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    vol = inference.query_volume(model, t, grid=grid, gather=False)
    b.record()
    sync_all()
    ms_vol = reduce_max([a.elapsed_time(b)])[0]
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        burst = float(peaks.get("bf16_tflops", 1590.0))
        rays = vps * img * img * world * args.steps
        flop = MLP_FWD_FLOP[(w["enc"], w["L"], w["H"])]
        vol_tflops = flop * n ** 3 / (ms_vol * 1e-3) * 1e-12 / world
        line = {"metric": "render_rays_per_s", "value": rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": args.workload, "detector": f"{img}x{img}", "novel_views_total": inf["views"],
                           "views_per_gpu_per_step": vps, "mlp": f"{w['L']}x{w['H']} {w['enc']}", "march_steps": 300, "grid": "128^3",
                           "trained_iterations": args.warmup + args.advance, "grid_occupied_fraction": occupied,
                           "volume_query": f"{n}^3 lattice on +-100, slabs sharded over the ranks",
                           "l2": "no flush: every step renders different views (1 M rays, > 100 MB of sample arrays per view)"},
                "spread": {"window_ms": window_ms},
                "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": vps * img * img * 4},
                "volume_query": {"points": n ** 3, "ms": ms_vol, "points_per_s": n ** 3 / (ms_vol * 1e-3),
                                 "roofline": {"bound": "tensor", "achieved": vol_tflops, "peak": burst, "unit": "TFLOP/s per GPU", "frac": vol_tflops / burst,
                                              "note": "includes building the lattice points with torch and the occupancy-grid lookup"}},
                "full_sweep_estimate_s": inf["views"] / (vps * world) * ms / args.steps * 1e-3,
                "gpu_launches": launches, "images_shape": list(imgs.shape), "volume_slab_shape": list(vol.shape)}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
