"""Inference throughput (BASELINE config 5 style): novel-view projections + attenuation-volume query through the fused
kernels (visualization.py:209-229, 315-354), views / volume slabs sharded across ranks.  Diagnostic companion of bench.py.

    python tools/bench_inference.py [--views 36] [--img 512] [--volume 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_inference.py
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerf_for_angiography_b200 as A  # noqa: E402
from nerf_for_angiography_b200 import inference  # noqa: E402
from nerf_for_angiography_b200.data import make_dataset  # noqa: E402
from nerf_for_angiography_b200.train import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=36)
    ap.add_argument("--img", type=int, default=512)
    ap.add_argument("--volume", type=int, default=256)
    ap.add_argument("--train-iters", type=int, default=300, help="brief training so that the occupancy grid is not trivially full")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    w = dict(img=64, thetas=[22.5 * i for i in range(8)], rays=4096, vol=128, L=4, H=128, enc="fourier")
    pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], kind="ct", volume_res=w["vol"], device=dev)
    torch.manual_seed(0)
    model = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev)
    tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"], lr=5e-4, seed=0, process_group=None)
    tr.world = 1                                                         # identical short training on every rank
    for _ in range(args.train_iters):
        tr.step()
    occ_frac = float(tr.acc_grid.binary.float().mean())
    views = [(360.0 * i / args.views, 0.0) for i in range(args.views)]
    src = np.array([0.0, 0.0, info["src_dist"]])
    kw = dict(views=views, src_pt=src, img_width=args.img, img_height=args.img, focal_length=7.5 * args.img, depth_samples_per_ray=300,
              near_thresh=info["near"], far_thresh=info["far"], early_stop_eps=1e-2, alpha_thre=1e-4)
    inference.render_projections(model, tr.acc_grid, tr.scene_aabb, **{**kw, "views": views[:max(world, 2)]})   # warm-up
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    if world > 1:
        torch.distributed.barrier()
    e0.record()
    imgs = inference.render_projections(model, tr.acc_grid, tr.scene_aabb, **kw)
    e1.record()
    t = torch.linspace(-100.0, 100.0, args.volume)
    vol = inference.query_volume(model, t, grid=tr.acc_grid)
    e2.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        n_rays = args.views * args.img * args.img
        print(json.dumps({"workload": "inference (config 5 style)", "n_gpus": world, "views": args.views, "detector": f"{args.img}x{args.img}",
                          "render_ms": float(ms[0]), "render_rays_per_s": n_rays / (float(ms[0]) * 1e-3),
                          "volume": f"{args.volume}^3", "volume_ms": float(ms[1]),
                          "volume_points_per_s": args.volume ** 3 / (float(ms[1]) * 1e-3), "grid_occupied_fraction": occ_frac,
                          "images_shape": list(imgs.shape), "volume_shape": list(vol.shape), "trained_iters": args.train_iters}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
