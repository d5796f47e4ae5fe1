"""Timing of the width-256 inference MLP forward (config 4, 8x256) on one large launch, default build next to variants selected by
environment variables:

    python tools/time_fwd256.py [--samples 4000000] [--reps 6] [--env NAME=VALUE ...]

Prints per variant the median CUDA-event time and the algorithmic TFLOP/s (1 065 984 FLOP/sample)."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--env", action="append", default=[], help="NAME=VALUE variant to time next to the default (repeatable)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    w = bench.WORKLOADS["config4"]
    torch.manual_seed(0)
    model = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev)
    model._ensure_flat()
    packed = A.ops.mlp_pack(model._desc, model._flat)
    R, n = 65536, args.samples
    g = torch.Generator(device=dev).manual_seed(1)
    o = torch.randn(R, 3, device=dev, generator=g) * 5 + torch.tensor([0.0, 0.0, 1500.0], device=dev)
    d = torch.randn(R, 3, device=dev, generator=g) * 0.05 + torch.tensor([0.0, 0.0, -1.0], device=dev)
    ri = torch.sort(torch.randint(0, R, (n,), device=dev, generator=g)).values.int()
    t0 = 1400.0 + torch.rand(n, device=dev, generator=g) * 199.0
    t1 = t0 + 2.0 / 3.0
    kw = dict(rays_o=o.contiguous(), rays_d=d.contiguous(), ray_idx=ri, t_starts=t0, t_ends=t1)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    flop = 1065984                                  # 8x256 Fourier, 2*MAC per sample
    ref = None
    for variant in [""] + args.env:
        name, _, val = variant.partition("=")
        if name:
            os.environ[name] = val
        out = A.ops.mlp_forward(model._desc, model._flat, packed, A.ops.OUT_ALPHA, A.ops.PREC_BF16, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = A.ops.mlp_forward(model._desc, model._flat, packed, A.ops.OUT_ALPHA, A.ops.PREC_BF16, **kw)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        if name:
            del os.environ[name]
        ms = float(np.median(ts))
        tf = flop * n / (ms * 1e-3) * 1e-12
        if ref is None:
            ref = out.clone()
        print(json.dumps({"variant": variant or "default", "samples": n, "ms_median": ms, "ms_min": min(ts), "tflops_algorithmic": tf,
                          "frac_of_burst_peak": tf / peaks.get("bf16_tflops", 1590.0), "same_as_default": bool(out.equal(ref))}))


if __name__ == "__main__":
    main()
