"""Host-side cost of Trainer.step() (time to ENQUEUE a step) next to the GPU time per step (diagnostic)."""
import os, sys, time, cProfile, pstats
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, nerf_for_angiography_b200 as A
from nerf_for_angiography_b200.data import make_dataset
from nerf_for_angiography_b200.train import Trainer
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["config3"]
pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], kind="ct", volume_res=w["vol"], device=dev)
model = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev)
tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"])
for _ in range(6): tr.step()
tr.n_iter = 257
torch.cuda.synchronize()
N = 14
t0 = time.perf_counter()
for _ in range(N): tr.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3 * (t1 - t0) / N:.3f} ms/step; enqueue + drain {1e3 * (t2 - t0) / N:.3f} ms/step")
if "--profile" in sys.argv:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(N): tr.step()
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
