import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nerf_for_angiography_b200.data import RayPool
V, H, W = 61, 512, 512
pool = RayPool(torch.eye(4, dtype=torch.float64, device="cuda").repeat(V, 1, 1), torch.rand(V, H, W, device="cuda"), 3840.0, None)
g = torch.Generator(device="cuda").manual_seed(1)
for w in (None, torch.rand(V, H, W, device="cuda") + 0.01):
    pool.weights = w; pool._wsum = None
    for _ in range(3): pool.sample(65536, generator=g)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): ids = pool.sample_ids(65536, generator=g)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(20): pool.gather(ids)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("weights" if w is not None else "uniform", f"sample_ids {(t1 - t0) / 20 * 1e3:.3f} ms  gather {(t2 - t1) / 20 * 1e3:.3f} ms  unique={ids.unique().numel()}")

