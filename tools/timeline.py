"""GPU timeline of a few training steps (torch.profiler / CUPTI): every kernel with its start offset, duration and the
idle gap before it, plus the busy fraction of the step.  Diagnostic only -- not a bench number.

    python tools/timeline.py [--workload config3] [--steps 2] [--start-iter 257]
    python tools/timeline.py --train-steps 70 --start-iter -1 --steps 2 --merge     # per-kernel totals after 70 real iterations
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerf_for_angiography_b200 as A  # noqa: E402
from nerf_for_angiography_b200.data import make_dataset  # noqa: E402
from nerf_for_angiography_b200.train import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--start-iter", type=int, default=257)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--train-steps", type=int, default=4, help="training steps before the captured ones (the field thins out as it trains)")
    ap.add_argument("--merge", action="store_true", help="print per-kernel totals instead of every launch")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:                                    # under torchrun: every rank trains, rank 0 prints its own timeline
        torch.distributed.init_process_group("nccl", device_id=dev)
    pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], kind="ct", volume_res=w["vol"], device=dev)
    model = A.CPPN(bench.model_def(w, dev, args.precision)).to(dev)
    tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"])
    for _ in range(args.train_steps):
        tr.step()
    if args.start_iter >= 0:
        tr.n_iter = args.start_iter
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            tr.step()
        torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA events captured")
        return
    t0 = evs[0].time_range.start
    end_prev = t0
    busy = 0.0
    print(f"{'start_us':>10} {'dur_us':>9} {'gap_us':>8}  kernel")
    merged = {}
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = e.time_range.start - end_prev
        busy += d
        end_prev = max(end_prev, e.time_range.end)
        if args.merge:
            m = merged.setdefault(e.name[:90], [0, 0.0])
            m[0] += 1; m[1] += d
        else:
            print(f"{s:10.1f} {d:9.1f} {gap:8.1f}  {e.name[:90]}")
    for name, (n, d) in sorted(merged.items(), key=lambda kv: -kv[1][1]):
        print(f"{d / args.steps:10.1f} us/step {n / args.steps:6.1f} launches/step  {name}")
    total = end_prev - t0
    print(f"steps={args.steps} span={total / 1e3:.3f} ms busy={busy / 1e3:.3f} ms ({100 * busy / total:.1f}%) per-step span={total / 1e3 / args.steps:.3f} ms")


if __name__ == "__main__":
    main()
