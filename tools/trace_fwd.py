"""Pipeline trace of the three-slot inference forward (CTA 0): per-stage hand-off / MMA issue / epilogue latencies in SM cycles.
    python tools/trace_fwd.py            # prints the event list of rounds 2..3 and a per-phase summary"""
import ctypes, os, sys
os.environ["ANGIO_TRACE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, nerf_for_angiography_b200 as A
from nerf_for_angiography_b200 import ops

dev = torch.device("cuda", 0)
w = bench.WORKLOADS["config3"]
model = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev); model._ensure_flat()
packed = ops.mlp_pack(model._desc, model._flat)
n = 128 * 148 * 3 * 8
x = ((torch.rand(n, 3, device=dev) * 2 - 1) * 100).contiguous()
ops.mlp_forward(model._desc, model._flat, packed, ops.OUT_SIGMA, ops.PREC_BF16, points=x); torch.cuda.synchronize()
buf = torch.zeros(8192, dtype=torch.int64, device=dev)
lib = ctypes.CDLL(A._lib.LIB_PATH); lib.angio_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.angio_debug_set_trace(buf.data_ptr()) == 0
ops.mlp_forward(model._desc, model._flat, packed, ops.OUT_SIGMA, ops.PREC_BF16, points=x); torch.cuda.synchronize()
lib.angio_debug_set_trace(None)
ev = buf.cpu().numpy()[:2048].reshape(16, 8, 4, 4)
names = {0: "MMA: a_ready seen", 1: "MMA: issued+commit", 2: "EPI: acc_ready seen", 3: "EPI: a_ready arrive"}
rec = [(int(ev[r, st, s, k]), k, s, st, r) for r in range(16) for st in range(8) for s in range(3) for k in range(4) if ev[r, st, s, k] != 0]
rec.sort()
t0 = rec[0][0]
for t, k, s, st, r in rec:
    if 2 <= r <= 3:
        print(f"{t - t0:8d}  round{r} slot{s} stage{st}  {names[k]}")
# per-phase summary over rounds 1..6, stages 1..3 (steady part of a tile)
issue, drain, epi, hand = [], [], [], []
for r in range(1, 7):
    for st in range(1, 4):
        for s in range(3):
            e = ev[r, st, s]
            nxt = ev[r, st + 1, s]
            if e.min() == 0 or nxt[0] == 0:
                continue
            issue.append(e[1] - e[0]); drain.append(e[2] - e[1]); epi.append(e[3] - e[2]); hand.append(nxt[0] - e[3])
for nm, v in (("issue 8 MMAs (a_ready seen -> commit)", issue), ("pipe drain + observe (commit -> acc_ready seen)", drain),
              ("epilogue (acc_ready seen -> a_ready arrive)", epi), ("hand-off (a_ready arrive -> issuer sees it, incl. its turn)", hand)):
    if v:
        print(f"{nm}: median {int(np.median(v))} cycles (min {int(min(v))}, max {int(max(v))}, n={len(v)})")
if issue:
    per_stage = np.median(issue) + np.median(drain) + np.median(epi) + np.median(hand)
    print(f"chain per stage and slot: {int(per_stage)} cycles; three slots x 512 MMA cycles / chain = {3 * 512 / per_stage:.2f} tensor-pipe bound")
