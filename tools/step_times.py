"""Per-step GPU time of the bench workload (CUDA events between steps), to see where the variance comes from (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, nerf_for_angiography_b200 as A
from nerf_for_angiography_b200.data import make_dataset
from nerf_for_angiography_b200.train import Trainer
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["config3"]
torch.manual_seed(0)
pool, info = make_dataset(img_size=w["img"], thetas=w["thetas"], kind="ct", volume_res=w["vol"], device=dev)
model = A.CPPN(bench.model_def(w, dev, "bf16")).to(dev)
tr = Trainer(model, pool, info["near"], info["far"], n_rays=w["rays"])
for _ in range(5): tr.step()
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
evs[0].record()
for i in range(40):
    tr.step(); evs[i + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(40)]
print(" ".join(f"{t:.2f}" for t in ts))
print(f"mean first 20: {sum(ts[:20]) / 20:.3f} ms, mean all: {sum(ts) / 40:.3f} ms, median: {sorted(ts)[20]:.3f}")
