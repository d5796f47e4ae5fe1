"""Time mlp_fwd_tc_kernel<ALPHA> on ray samples: direct vs through an index list, at the visibility-pass sizes (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, nerf_for_angiography_b200 as A
from nerf_for_angiography_b200 import ops
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["config3"]
model = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev); model._ensure_flat()
packed = ops.mlp_pack(model._desc, model._flat)
R, per = 65536, 260
n = R * per
o = torch.tensor([0.0, 0.0, 1500.0], device=dev).repeat(R, 1).contiguous()
d = torch.nn.functional.normalize(torch.randn(R, 3, device=dev) * 0.03 + torch.tensor([0, 0, -1.0], device=dev), dim=1).contiguous()
ray_idx = torch.arange(R, device=dev, dtype=torch.int32).repeat_interleave(per).contiguous()
t0 = (1400.0 + (torch.arange(per, device=dev) * (200.0 / 300)).repeat(R)).contiguous()
t1 = (t0 + 200.0 / 300).contiguous()
kw = dict(rays_o=o, rays_d=d, ray_idx=ray_idx, t_starts=t0, t_ends=t1)
head = (torch.arange(R, device=dev, dtype=torch.int32)[:, None] * per + torch.arange(32, device=dev, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
out = torch.empty(n, device=dev)

def timed(label, fn, n_eval):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label:46s} {ms:7.3f} ms  {n_eval / ms * 1e-3:6.2f} M samples/ms  {139776 * n_eval / ms * 1e-9:5.0f} TFLOP/s")

timed("all 17 M samples, direct", lambda: ops.mlp_forward(model._desc, model._flat, packed, ops.OUT_ALPHA, ops.PREC_BF16, out=out, **kw), n)
m = head.numel()
kw_small = dict(rays_o=o, rays_d=d, ray_idx=ray_idx[:m].contiguous(), t_starts=t0[:m].contiguous(), t_ends=t1[:m].contiguous())
timed("first 2.1 M samples, direct (contiguous arrays)", lambda: ops.mlp_forward(model._desc, model._flat, packed, ops.OUT_ALPHA, ops.PREC_BF16, **kw_small), m)
timed("32 samples/ray through sample_idx (2.1 M)", lambda: ops.mlp_forward(model._desc, model._flat, packed, ops.OUT_ALPHA, ops.PREC_BF16, out=out, sample_idx=head, **kw), m)
