"""Per-step GPU time of the bench workload (CUDA events between steps), to see where the variance comes from (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, nerf_for_angiography_b200 as A
from nerf_for_angiography_b200.data import make_dataset
from nerf_for_angiography_b200.train import Trainer
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
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
import time
time.sleep(0.3 * rank)
print(f"rank {rank} peer={getattr(tr, 'peer', None) is not None}:", " ".join(f"{t:.2f}" for t in ts[:24]), f"| mean first 20: {sum(ts[:20]) / 20:.3f} ms", flush=True)
if world > 1:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
