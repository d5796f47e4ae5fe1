"""Reconstruction-quality check: train the same synthetic phantom with the bf16 tcgen05 path and the fp32 check path
(reference arithmetic) from identical weights / ray draws and compare the test-view PSNR (run_nerf_acc.py:338-357).

    python tools/psnr_check.py [--iters 2000] [--img 64] [--rays 4096]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerf_for_angiography_b200 as A  # noqa: E402
from nerf_for_angiography_b200.data import make_dataset  # noqa: E402
from nerf_for_angiography_b200.train import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--img", type=int, default=64)
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--lr", type=float, default=5e-4)
    ap.add_argument("--every", type=int, default=250, help="evaluate the test view every this many iterations")
    ap.add_argument("--workload", default=None, help="take detector size, views and phantom from bench.WORKLOADS (e.g. config2)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    w = dict(img=args.img, thetas=[22.5 * i for i in range(8)], rays=args.rays, vol=128, L=4, H=128, enc="fourier", kind="ct")
    if args.workload:
        w = dict(bench.WORKLOADS[args.workload], rays=args.rays)
    pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], kind=w["kind"], volume_res=w["vol"], device=dev, weight_strategy="distance")
    print(f"workload: {w['img']}x{w['img']} detector, {len(w['thetas']) + 1} views, phantom {w['kind']}, {w['rays']} rays/iteration, lr {args.lr}")
    res = {}
    for prec in ("fp32", "bf16"):
        torch.manual_seed(0)
        model = A.CPPN(bench.model_def(w, dev, prec)).to(dev)
        tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"], lr=args.lr, seed=0)
        t0 = time.perf_counter()
        curve = []
        for it in range(args.iters):
            out = tr.step()
            if (it + 1) % args.every == 0:
                ev = tr.evaluate()
                curve.append((it + 1, round(ev["psnr"], 3), round(float(out["loss"]), 6)))
        torch.cuda.synchronize()
        ev = tr.evaluate()
        res[prec] = dict(psnr=ev["psnr"], mse=ev["mse"], seconds=time.perf_counter() - t0, curve=curve,
                         kept_samples_last=out["n_samples"], marched_last=out["n_samples_prefilter"])
        print(prec, json.dumps(res[prec]))
    d = abs(res["bf16"]["psnr"] - res["fp32"]["psnr"])
    print(f"test-view PSNR after {args.iters} iterations: fp32 {res['fp32']['psnr']:.2f} dB, bf16 {res['bf16']['psnr']:.2f} dB (|diff| {d:.2f} dB)")
    # a single evaluation is noisy at this learning rate (the curves oscillate by several dB): compare the second half of the curves
    import statistics
    for name, f in (("median", statistics.median), ("best", max)):
        a, b = (f([c[1] for c in res[k]["curve"] if c[0] > args.iters // 2]) for k in ("fp32", "bf16"))
        print(f"{name} test-view PSNR over iterations {args.iters // 2}..{args.iters}: fp32 {a:.2f} dB, bf16 {b:.2f} dB (bf16 - fp32 = {b - a:+.2f} dB)")


if __name__ == "__main__":
    main()
