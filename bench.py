#!/usr/bin/env python
"""bench.py -- training-throughput benchmark of the NeRF-for-angiography hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config3|config2|tiny]

One "step" = one reference training iteration (/root/reference/nerf/run_nerf_acc.py:263-328): sample a ray batch,
refresh the occupancy grids (every 16th step), march + visibility filter, MLP forward, Beer-Lambert composite + MSE,
backward, Adam.  Default workload: BASELINE.json configs[2] ("config3": 512x512 cone-beam, 60 views + test view,
4x128 Fourier MLP, occupancy-grid marching, 65 536 rays per GPU per step), synthetic phantom, random-init weights.

Prints ONE JSON line (rank 0).  `value` = rays/s with the step's inputs resident in HBM; `e2e` = the same metric with
each step's ray batch arriving from pinned HOST memory (the reference samples rays on the host and copies them every
iteration, nerf/nerf_helpers.py:144-148) and the loss read back; `roofline` describes the dominant kernel (the fused
tcgen05 MLP forward of the no-grad visibility pass); `cpu_baseline` is the oracle port timed on this box's host cores.
"""
import argparse
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
    # name: (img_size, n_views(theta sweep), rays/GPU/step, volume_res, hidden layers, width, pos_enc)
    "config3": dict(img=512, thetas=[6.0 * i for i in range(60)], rays=65536, vol=256, L=4, H=128, enc="fourier"),
    "config2": dict(img=256, thetas=[0.0, 45.0, 90.0, 135.0], rays=65536, vol=256, L=4, H=128, enc="fourier"),
    # BASELINE configs[3]: 1024^2 x 120 views, 8x256 MLP.  Width 256 has no tcgen05 kernel yet: fp32 path.
    "config4": dict(img=1024, thetas=[3.0 * i for i in range(120)], rays=131072, vol=256, L=8, H=256, enc="fourier", precision="fp32"),
    "tiny": dict(img=64, thetas=[22.5 * i for i in range(8)], rays=4096, vol=64, L=4, H=128, enc="fourier"),
}
MLP_FWD_FLOP = {("fourier", 4, 128): 139776, ("none", 4, 128): 132096, ("fourier", 8, 256): 1065984}


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
    n_cpu = 8192 if args.workload != "tiny" else 256                  # ~1.5 s of CPU work per step on 16 cores
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_step_time(w, n_cpu, 1)
    t, n_pre, n_kept = cpu_oracle_step_time(w, n_cpu, max(1, min(args.steps, 8)))
    val = n_cpu / t
    line = {"impl": "reference", "metric": "train_rays_per_s", "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sample": f"{n_cpu} rays/step of view 0, full 128^3 grid, {n_pre} marched / {n_kept} kept samples"},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n_cpu}-ray training steps (oracle port: torch-CPU fp32 MLP + C marcher)"},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
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

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import nerf_for_angiography_b200 as A
    from nerf_for_angiography_b200.data import make_dataset
    from nerf_for_angiography_b200.train import Trainer
    lib = A._lib.load()

    torch.manual_seed(0)
    pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], test_view=(135.0, 135.0), kind="ct", volume_res=w["vol"],
                              device=dev, seed=0)
    model = A.CPPN(model_def(w, dev, args.precision)).to(dev)          # same seed => identical weights on every rank
    tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"], seed=0, early_termination=0 if args.full_visibility else 32)
    R = w["rays"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    # ---------------- e2e batches: pinned host memory, one per step (reference: host-side sampling + H2D every iteration)
    host_batches = []
    for _ in range(args.steps + args.warmup):
        o, d, t = pool.sample(R, generator=tr.ray_gen)
        host_batches.append(tuple(x.cpu().pin_memory() for x in (o, d, t)))
    torch.cuda.synchronize()

    # ---------------- device-resident arm
    clocks = ClockSampler(local_rank)
    if not args.no_clocks:
        clocks.prepare()
    for _ in range(args.warmup):
        tr.step()
    sync_all()
    snap = tr.snapshot()                                               # both arms are timed from this model / grid / optimiser state
    # Rehearsal (untimed): run the exact sequence that is about to be timed once, then rewind.  The step sizes are
    # deterministic, so afterwards the caching allocator owns every block the timed region needs (the snapshot's clones and
    # the occupancy-grid refresh at iteration 16 otherwise trigger cudaMalloc calls inside it -- usually 5 ms each, sometimes
    # 100-250 ms with 55 GB already reserved -- during which the stream drains).
    for _ in range(args.steps):
        tr.step()
    sync_all()
    tr.restore(snap)
    launches0 = int(lib.angio_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # The steps are enqueued from Python with the GPU 5x slower than the host; a generational GC pass (100-200 ms with torch,
    # numpy and pandas loaded) in the middle of the timed region would drain the stream and be billed to the kernels.
    gc.collect()
    gc.disable()
    tr.kernel_events = []                                              # (start, end, sample count) per visibility-pass MLP launch
    step_totals = []                                                   # device counters of every step, read after the timed region
    torch.cuda.profiler.start()                                        # ncu --profile-from-start off captures exactly the timed region
    e0.record()
    step_events, host_dbg, host_t = [], [], time.perf_counter()
    every = max(1, args.steps // 8)
    for i_step in range(args.steps):
        out = tr.step()
        step_totals.append(out["totals"])
        if not args.no_clocks and i_step % every == every - 1:
            clocks.sample()                                            # the GPU is executing the steps enqueued so far
        if os.environ.get("BENCH_DEBUG"):
            ev = torch.cuda.Event(enable_timing=True); ev.record(); step_events.append(ev)
            now = time.perf_counter()
            st = torch.cuda.memory_stats()
            host_dbg.append((now - host_t, st["num_device_alloc"], st["num_device_free"], st["num_alloc_retries"]))
            host_t = now
    e1.record()
    sync_all()
    gc.enable()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    if step_events:
        prev, ts = e0, []
        for ev in step_events:
            ts.append(prev.elapsed_time(ev)); prev = ev
        print(f"[rank {rank}] per-step ms: " + " ".join(f"{t:.2f}" for t in ts), file=sys.stderr, flush=True)
        print(f"[rank {rank}] host ms/step (device allocs, frees, retries): " +
              " ".join(f"{1e3 * h:.1f}({a},{f},{r})" for h, a, f, r in host_dbg), file=sys.stderr, flush=True)
    clk = clocks.result() if not args.no_clocks else None
    launches = int(lib.angio_launch_count()) - launches0
    host_totals = [t.tolist() if isinstance(t, torch.Tensor) else list(t) for t in step_totals]
    if any(len(t) > 3 and t[3] != 0 for t in host_totals):
        raise RuntimeError("ray sampler overflow during the timed region")
    n_pre_total = sum(t[0] for t in host_totals)
    n_kept_total = sum(t[1] for t in host_totals)
    kernel_ms = [a.elapsed_time(b) for a, b, _ in tr.kernel_events]
    kernel_n = [int(c.item()) if isinstance(c, torch.Tensor) else int(c) for _, _, c in tr.kernel_events]   # samples each launch evaluated
    tr.kernel_events = None
    last_loss = float(out["loss"])

    # ---------------- e2e arm: host buffers in, loss out, every step (same starting state as the device-resident arm)
    tr.restore(snap)
    for i in range(args.warmup):
        o, d, t = (x.to(dev, non_blocking=True) for x in host_batches[i])
        float(tr.step(rays=(o, d, t))["loss"])
    tr.restore(snap)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()
    e2.record()
    for i in range(args.steps):
        o, d, t = (x.to(dev, non_blocking=True) for x in host_batches[args.warmup + i])
        loss_host = float(tr.step(rays=(o, d, t))["loss"])            # D2H read of the step's loss
    e3.record()
    sync_all()
    gc.enable()
    ms_e2e = e2.elapsed_time(e3)

    t_ms = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    cnt = torch.tensor([n_pre_total, n_kept_total], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(cnt, op=torch.distributed.ReduceOp.SUM)
    ms, ms_e2e = float(t_ms[0]), float(t_ms[1])
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        flop = MLP_FWD_FLOP.get((w["enc"], w["L"], w["H"]))
        roofline = None
        if kernel_ms and flop and args.precision == "bf16":
            # every entry is one launch of the visibility-pass MLP forward (two per step with early ray termination: the first
            # 32 samples of every ray, then the rest of the rays still alive -- usually none at random init).  Launches that
            # do not fill the 148 SMs twice are left out; the rest are weighted by the samples they evaluated.
            big = [(n, t) for n, t in zip(kernel_n, kernel_ms) if t > 0 and n >= 2 * 128 * 148]
            if big:
                ach = flop * sum(n for n, _ in big) / (sum(t for _, t in big) * 1e-3) * 1e-12
                peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
                roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                            "kernel": "mlp_fwd_tc_kernel<ALPHA> (no-grad visibility pass)",
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                            "avg_launch_ms": float(np.mean([t for _, t in big])), "samples_per_launch": float(np.mean([n for n, _ in big])),
                            "launches_timed": len(big)}
                try:   # DRAM bytes of this kernel from the committed `ncu --set full` capture, scaled to this run's samples per launch
                    tr_ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["mlp_fwd_tc_kernel<ALPHA>"]
                    roofline["traffic"] = tr_ncu["dram_bytes"] / tr_ncu["samples"] * roofline["samples_per_launch"]
                    roofline["traffic_source"] = tr_ncu["source"]
                except Exception:
                    pass
        cpu = None
        if not args.no_cpu_baseline:
            n_cpu, n_cpu_steps = (8192, 6) if args.workload != "tiny" else (256, 2)   # ~10 s of CPU work on the box's host cores
            cpu_oracle_step_time(w, n_cpu, 1)                                          # untimed warm-up (thread pools, allocator)
            t_cpu, cp, ck = cpu_oracle_step_time(w, n_cpu, n_cpu_steps)
            cpu = {"value": n_cpu / t_cpu, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{n_cpu_steps} training steps of {n_cpu} rays of the oracle port ({cp} marched / {ck} kept samples per step, "
                             f"{t_cpu:.2f} s per step)"}
        rays = R * world * args.steps
        line = {"metric": "train_rays_per_s", "value": rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "detector": f"{w['img']}x{w['img']}", "views": len(w["thetas"]) + 1,
                           "mlp": f"{w['L']}x{w['H']} {w['enc']}", "rays_per_gpu_per_step": R, "march_steps": 300, "grid": "128^3",
                           "iterations": [args.warmup, args.warmup + args.steps],
                           "l2": "no flush: every step draws fresh rays and streams the saved bf16 tile images of its kept samples "
                                 f"(~1 KB/sample x {n_kept_total / max(args.steps, 1) / 1e6:.1f} M samples/step here) plus the sample arrays "
                                 "through HBM, far more than the 126 MB L2",
                           "visibility_pass": "every marched sample (reference order)" if args.full_visibility else
                                              "early ray termination: first 32 samples of every ray, then only the rays still transparent "
                                              + ("(marched lazily as well) " if tr.lazy_march else "") +
                                              "-- kept samples bit-identical to evaluating every sample"},
                "mlp_evals_visibility_per_step": float(sum(kernel_n)) / max(args.steps, 1),
                "samples_per_s_marched": float(cnt[0]) / (ms * 1e-3), "samples_per_s_kept": float(cnt[1]) / (ms * 1e-3),
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": R * 28, "d2h_bytes_per_step": 4 + 8},
                "roofline": roofline, "cpu_baseline": cpu, "final_loss": last_loss, "e2e_final_loss": loss_host}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
