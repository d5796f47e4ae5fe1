"""Run-to-run determinism of the training loop: the same seed trained twice must give bit-identical parameters, with and without
evaluations of the test view in between (evaluate() must not perturb the training state).

    python tools/check_determinism.py [--workload config2] [--rays 5625] [--iters 600] [--precision bf16]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerf_for_angiography_b200 as A  # noqa: E402
from nerf_for_angiography_b200.data import make_dataset  # noqa: E402
from nerf_for_angiography_b200.train import Trainer  # noqa: E402


def run(w, pool, info, prec, iters, eval_every, dev, lr):
    torch.manual_seed(0)
    model = A.CPPN(bench.model_def(w, dev, prec)).to(dev)
    tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"], lr=lr, seed=0)
    trace = []
    for it in range(iters):
        out = tr.step()
        if eval_every and (it + 1) % eval_every == 0:
            tr.evaluate()
        if (it + 1) % 50 == 0:
            trace.append((it + 1, tr.flat.detach().clone(), float(out["loss"])))
    torch.cuda.synchronize()
    return trace


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--rays", type=int, default=5625)
    ap.add_argument("--iters", type=int, default=600)
    ap.add_argument("--lr", type=float, default=5e-4)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    w = dict(bench.WORKLOADS[args.workload], rays=args.rays)
    pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], kind=w["kind"], volume_res=w["vol"], device=dev, weight_strategy="distance")
    a = run(w, pool, info, args.precision, args.iters, 0, dev, args.lr)
    ok = True
    for name, ev in (("repeat", 0), ("with evaluate() every 100", 100)):
        b = run(w, pool, info, args.precision, args.iters, ev, dev, args.lr)
        first = next((ia for (ia, pa, _), (_, pb, _) in zip(a, b) if not torch.equal(pa, pb)), None)
        if first is None:
            print(f"{args.precision} {name}: parameters bit-identical at every 50th of {args.iters} iterations")
        else:
            ok = False
            (_, pa, la), (_, pb, lb) = next((x, y) for x, y in zip(a, b) if x[0] == first)
            print(f"{args.precision} {name}: FIRST DIFFERENCE at iteration <= {first}: max |dp| {float((pa - pb).abs().max()):.3e}, loss {la:.6g} vs {lb:.6g}")
    print("DETERMINISTIC" if ok else "NOT DETERMINISTIC")


if __name__ == "__main__":
    main()
