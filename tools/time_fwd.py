"""A/B timing of the inference MLP forward (visibility pass): two-slot vs three-slot tcgen05 kernel on one large launch.

    python tools/time_fwd.py [--samples 7000000] [--reps 10]

Prints per kernel the median CUDA-event time, algorithmic TFLOP/s (139 776 FLOP/sample, 4x128 Fourier) and the fraction of the
measured bf16 peaks, and checks that both kernels return bit-identical outputs."""
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
    ap.add_argument("--samples", type=int, default=7_000_000)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda")
    w = bench.WORKLOADS["config3"]
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
    outs = {}
    for slots in ("2", "3", "3s"):                 # "3": three slots in lock step, "3s": staggered over the layers (the default)
        os.environ["ANGIO_FWD_SLOTS"] = slots[0]
        os.environ["ANGIO_FWD_STAGGER"] = "1" if slots.endswith("s") else "0"
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
        ms = float(np.median(ts))
        tf = 139776 * n / (ms * 1e-3) * 1e-12
        outs[slots] = out.clone()
        print(json.dumps({"slots": slots, "samples": n, "ms_median": ms, "ms_min": min(ts), "tflops_algorithmic": tf,
                          "frac_of_burst_peak": tf / peaks.get("bf16_tflops", 1590.0), "frac_of_sustained_peak": tf / peaks.get("bf16_tflops_sustained", 1400.0)}))
    same = bool(outs["2"].equal(outs["3"])) and bool(outs["2"].equal(outs["3s"]))
    print("bit-identical outputs:", same, " max |diff|:", float((outs["2"] - outs["3"]).abs().max()), float((outs["2"] - outs["3s"]).abs().max()))
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
