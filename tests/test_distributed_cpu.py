"""CPU-only, world_size = 2 over gloo: the host-side logic of the multi-GPU path (sharding, ragged gather, and that the
all-reduced per-rank gradients equal the 1-rank gradient on the concatenated batch)."""
import functools
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from nerf_for_angiography_b200.distributed import allreduce_mean_gradient, gather_concat, shard_range, world
    from oracle import cppn as ocppn, geometry as ogeo, nerfacc_ref, pipeline
    assert world() == (rank, world_size)
    # ---- ragged gather
    local = torch.arange(3 + rank, dtype=torch.float32)[:, None] + 100 * rank
    got = gather_concat(local)
    if rank == 0:
        assert got.shape[0] == sum(3 + r for r in range(world_size))
        assert torch.equal(got[:3, 0], torch.arange(3.0)) and float(got[3, 0]) == 100.0
    else:
        assert got is None
    # ---- sharded evaluation (the occupancy-grid refresh of data-parallel training): every rank evaluates 1/world of the rows and
    #      all-gathers -- identical to evaluating everything, for row counts that do and do not divide by the world size
    from nerf_for_angiography_b200.distributed import evaluate_sharded
    for n in (1, 7, 64, 1001):
        x = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * 0.01
        calls = []
        fn = lambda xs: (calls.append(xs.shape[0]), torch.sin(xs).sum(-1))[1]          # noqa: E731
        got = evaluate_sharded(fn, x, rank, world_size)
        assert torch.equal(got, torch.sin(x).sum(-1)) and got.shape == (n,)
        assert sum(calls) <= (n + world_size - 1) // world_size                        # this rank only did its share
        buf = torch.full((2 * ((n + 1) // 2) + 5,), -1.0)
        assert torch.equal(evaluate_sharded(fn, x, rank, world_size, out=buf), got)
    # ---- data-parallel gradient == single-process gradient on the concatenated batch (oracle math, fp32)
    roi = np.array([-100, -100, -100, 100, 100, 100], np.float32)
    grid = nerfacc_ref.OccupancyGrid(roi, 16); grid.binary[:] = True; grid.occs[:] = 0.05
    p = ocppn.init_params(2, 64, "fourier", 5, 0.05, seed=0)
    p["output_linear.0.bias"] = p["output_linear.0.bias"] - 3.0
    o, d, _ = ogeo.get_ray_values(20.0, 0.0, 0.0, [0, 0, 1500.0], 12, 12, 90.0)
    o = o.reshape(-1, 3).astype(np.float32); d = d.reshape(-1, 3).astype(np.float32)
    target = torch.from_numpy(np.random.default_rng(0).random(len(o)).astype(np.float32))

    def flat_grad(idx):
        params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        f = functools.partial(ocppn.cppn_forward, params, pos_enc="fourier", basis=5)
        pix, _ = pipeline.render_rays(f, grid, roi, o[idx], d[idx], 100, 1400.0, 1600.0, 1e-2, 1e-4)
        torch.nn.functional.mse_loss(pix, target[idx]).backward()
        return torch.cat([params[k].grad.reshape(-1) for k in sorted(params)])

    n = len(o)
    lo, hi = shard_range(n, rank, world_size)
    g = flat_grad(np.arange(lo, hi))
    total = allreduce_mean_gradient(g, hi - lo)
    assert total == n
    g_ref = flat_grad(np.arange(n))
    assert torch.allclose(g, g_ref, rtol=1e-4, atol=1e-9), float((g - g_ref).abs().max())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    sys.path.insert(0, ROOT)
    from nerf_for_angiography_b200.distributed import shard_range
    for n in (0, 1, 7, 360, 512):
        for ws in (1, 2, 3, 8):
            parts = [shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
