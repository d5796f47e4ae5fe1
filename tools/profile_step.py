"""Per-phase timing of one training step (synchronising between phases; wall-clock per phase, not a bench number).

    python tools/profile_step.py [--workload config3] [--steps 5] [--start-iter 256]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerf_for_angiography_b200 as A  # noqa: E402
from nerf_for_angiography_b200 import ops  # noqa: E402
from nerf_for_angiography_b200.data import make_dataset  # noqa: E402
from nerf_for_angiography_b200.train import Trainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--start-iter", type=int, default=256)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    w = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], kind="ct", volume_res=w["vol"], device=dev)
    model = A.CPPN(bench.model_def(w, dev, args.precision)).to(dev)
    tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"])
    for _ in range(3):
        tr.step()
    tr.n_iter = args.start_iter
    acc = {}

    def timed(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    m = tr.model
    for _ in range(args.steps):
        o, d, target = timed("sample_rays", lambda: pool.sample(tr.n_rays, generator=tr.ray_gen))
        timed("pack_weights", tr._refresh_packed)
        timed("grid_update", tr.update_grids)
        g = tr.acc_grid
        ray_idx, t0, t1, offsets = timed("march(count+scan+write)", lambda: ops.march(o, d, tr.scene_aabb, g._roi_host, g._resolution,
                                                                                   g._binary_u8(), tr.near, tr.far, tr.step_size))
        kw = dict(rays_o=o, rays_d=d, ray_idx=ray_idx, t_starts=t0, t_ends=t1)
        alphas = timed("mlp_alpha(visibility pass)", lambda: ops.mlp_forward(m._desc, tr.flat, tr.packed, ops.OUT_ALPHA, m._precision_id, **kw))
        thre = min(tr.alpha_thre, g.occs_mean_host)
        ray_idx2, t02, t12, off2, _ = timed("visibility+compact", lambda: ops.visibility_compact(alphas, offsets, t0, t1, tr.early_stop_eps, thre))
        kw2 = dict(rays_o=o, rays_d=d, ray_idx=ray_idx2, t_starts=t02, t_ends=t12)
        logits, saved = timed("mlp_forward(train)", lambda: ops.mlp_forward(m._desc, tr.flat, tr.packed, ops.OUT_LOGIT, m._precision_id, saved=True, pool=tr.pool_bufs, **kw2))
        pix, gl, loss = timed("composite+mse+bwd", lambda: ops.composite_mse_fused(logits, t02, t12, off2, target, tr.n_rays))
        timed("mlp_backward", lambda: ops.mlp_backward(m._desc, tr.flat, tr.packed, saved, gl, m._precision_id, grad_params=tr.grad, pool=tr.pool_bufs, **kw2))
        timed("adam", lambda: ops.adam_step(tr.flat, tr.grad, tr.exp_avg, tr.exp_avg_sq, tr.lr, tr.n_iter_adam + 1))
        tr.n_iter_adam += 1
        tr.n_iter += 1
        n_pre, n_kept = ray_idx.numel(), ray_idx2.numel()
    tot = sum(acc.values())
    print(f"workload={args.workload} rays={tr.n_rays} marched={n_pre} kept={n_kept} loss={float(loss) / tr.n_rays:.5f}")
    for k, v in acc.items():
        print(f"  {k:32s} {v / args.steps:9.3f} ms/step  {100 * v / tot:5.1f}%")
    print(f"  {'TOTAL':32s} {tot / args.steps:9.3f} ms/step")


if __name__ == "__main__":
    main()
