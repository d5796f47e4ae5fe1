"""Pipeline trace of mlp_fwd_tc_kernel (CTA 0): per-stage handoff / MMA / epilogue latencies in SM cycles."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, nerf_for_angiography_b200 as A
from nerf_for_angiography_b200 import ops

dev = torch.device("cuda", 0)
w = bench.WORKLOADS["config3"]
model = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev); model._ensure_flat()
packed = ops.mlp_pack(model._desc, model._flat)
n = 128 * 148 * 8
x = ((torch.rand(n, 3, device=dev) * 2 - 1) * 100).contiguous()
ops.mlp_forward(model._desc, model._flat, packed, ops.OUT_SIGMA, ops.PREC_BF16, points=x); torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib = ctypes.CDLL(A._lib.LIB_PATH); lib.angio_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.angio_debug_set_trace(buf.data_ptr()) == 0
ops.mlp_forward(model._desc, model._flat, packed, ops.OUT_SIGMA, ops.PREC_BF16, points=x); torch.cuda.synchronize()
lib.angio_debug_set_trace(None)
ev = buf.cpu().numpy()[:1024].reshape(16, 8, 2, 4)
names = {0: "MMA: a_ready seen", 1: "MMA: issued+commit", 2: "EPI: acc_ready seen", 3: "EPI: a_ready arrive"}
rec = [(int(ev[r, st, s, k]), k, s, st, r) for r in range(16) for st in range(8) for s in range(2) for k in range(4) if ev[r, st, s, k] != 0]
rec.sort()
t0 = rec[0][0]
for t, k, s, st, r in rec:
    if 2 <= r <= 4:
        nm = names[k] if st < 6 else {2: "EPI: acc drained, next tile signalled", 3: "EPI: last-layer math done"}[k]
        print(f"{t - t0:8d}  round{r} slot{s} stage{st}  {nm}")
