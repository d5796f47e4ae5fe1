"""Multi-GPU correctness on hardware (skipped below 2 GPUs): an N-rank data-parallel step equals the 1-rank step on the concatenated
batch, the sharded occupancy-grid refresh leaves bit-identical grids on every rank, replicas stay bit-identical -- through the
fused NVLink peer-memory all-reduce + Adam kernel (angio_adam_step_allreduce) and through the NCCL all_reduce path (ANGIO_P2P=0).
tools/check_dp_equivalence.py does the work under torchrun; its logs for N = 2 and 8 are kept under profiles/."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_data_parallel_step_equals_single_rank(p2p):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    env = dict(os.environ, ANGIO_P2P=p2p)
    port = 29500 + (os.getpid() % 400) + (7 if p2p == "0" else 0)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "check_dp_equivalence.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DP EQUIVALENCE OK" in r.stdout
    assert ("NVLink peer memory" in r.stdout) == (p2p == "1"), r.stdout[-2000:]
