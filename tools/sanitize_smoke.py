"""Tiny invocations of the kernels added in round 2, meant to run under `compute-sanitizer --tool memcheck`:
three-slot forward (plain / device count / index list), width-256 forward + training forward + backward, thread-per-ray
visibility mask, lazy march + filter, a failed sampler draw."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nerf_for_angiography_b200 as A  # noqa: E402
from nerf_for_angiography_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
R, n = 64, 1000
g = torch.Generator(device=dev).manual_seed(1)
o = (torch.randn(R, 3, device=dev, generator=g) * 5 + torch.tensor([0.0, 0.0, 1500.0], device=dev)).contiguous()
d = (torch.randn(R, 3, device=dev, generator=g) * 0.05 + torch.tensor([0.0, 0.0, -1.0], device=dev)).contiguous()
ri = torch.sort(torch.randint(0, R, (n,), device=dev, generator=g)).values.int()
t0 = 1400.0 + torch.rand(n, device=dev, generator=g) * 199.0
t1 = t0 + 2.0 / 3.0
kw = dict(rays_o=o, rays_d=d, ray_idx=ri, t_starts=t0, t_ends=t1)
for L, H in ((4, 128), (2, 256)):
    w = dict(L=L, H=H, enc="fourier")
    m = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev)
    m._ensure_flat()
    packed = ops.mlp_pack(m._desc, m._flat)
    a = ops.mlp_forward(m._desc, m._flat, packed, ops.OUT_ALPHA, ops.PREC_BF16, **kw)
    nd = torch.tensor([333], dtype=torch.int32, device=dev)
    b = ops.mlp_forward(m._desc, m._flat, packed, ops.OUT_ALPHA, ops.PREC_BF16, n_dev=nd, **kw)
    ids = torch.arange(1, n, 5, dtype=torch.int32, device=dev)
    out = torch.zeros(n, device=dev)
    ops.mlp_forward(m._desc, m._flat, packed, ops.OUT_ALPHA, ops.PREC_BF16, out=out, sample_idx=ids, **kw)
    y, saved = ops.mlp_forward(m._desc, m._flat, packed, ops.OUT_LOGIT, ops.PREC_BF16, saved=True, **kw)
    gr = ops.mlp_backward(m._desc, m._flat, packed, saved, torch.randn(n, device=dev), ops.PREC_BF16, **kw)
    torch.cuda.synchronize()
    assert bool(b[:333].equal(a[:333])) and bool(out[ids.long()].equal(a[ids.long()])) and bool(torch.isfinite(gr).all())
    print(f"{L}x{H}: forward / n_dev / index list / training forward / backward OK")
# lazy march + filter (thread-per-ray visibility mask on the tail) on a small grid
w = dict(L=4, H=128, enc="fourier")
m = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev)
m._ensure_flat()
with torch.no_grad():
    m.output_linear[0].bias -= 6.0
packed = ops.mlp_pack(m._desc, m._flat)
roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
binary = torch.ones((32, 32, 32), dtype=torch.bool, device=dev)
oo = torch.tensor([[0.0, 0.0, 1500.0]], device=dev).repeat(R, 1).contiguous()
dd = torch.cat([torch.rand(R, 2, device=dev) * 0.1 - 0.05, -torch.ones(R, 1, device=dev)], 1).contiguous()
totals = torch.zeros(4, dtype=torch.int32, device=dev)
res = ops.march_filter_lazy(m._desc, m._flat, packed, ops.PREC_BF16, oo, dd, roi, roi, 32, binary, 1400.0, 1600.0, 200.0 / 300, 1e-2, 1e-4,
                            k0=32, totals=totals, thre_cap=torch.tensor([0.5], device=dev))
torch.cuda.synchronize()
print("lazy march + filter OK:", totals.tolist())
wts = torch.zeros(4096, device=dev); wts[:10] = 1.0
ids, status = ops.sample_without_replacement(100, 4096, wts, 10.0, 10.0, 7, dev)
torch.cuda.synchronize()
assert status.tolist()[1] == 1 and int(ids.max()) == 0
print("failed sampler draw leaves valid ids OK")
