import sys, torch
sys.path.insert(0, "/root/repo")
import nerf_for_angiography_b200 as A
from nerf_for_angiography_b200 import ops
for n in (1024, 32768, 65536, 262144, 1048576):
    c = torch.randint(0, 300, (n,), dtype=torch.int32, device="cuda")
    for _ in range(3): o = ops.exclusive_scan(c)
    assert int(o[-1]) == int(c.sum())
    assert o[:-1].equal((torch.cumsum(c, 0) - c).int())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): ops.exclusive_scan(c)
    e1.record(); torch.cuda.synchronize()
    print(n, f"{e0.elapsed_time(e1) / 50 * 1e3:.1f} us per scan (incl. launch + alloc)")
