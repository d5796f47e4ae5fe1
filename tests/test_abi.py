"""CPU-only: the C-ABI shared library loads and exports every symbol include/angio_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "angio_b200.h")).read()
    return sorted(set(re.findall(r"ANGIO_API\s+[\w\s\*]+?\b(angio_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared_symbols()
    for must in ["angio_raygen", "angio_march_count", "angio_march_write", "angio_visibility_mask", "angio_compact_samples",
                 "angio_mlp_forward", "angio_mlp_backward", "angio_composite_forward", "angio_composite_backward",
                 "angio_grid_ema_update", "angio_adam_step", "angio_last_error_string", "angio_version"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from nerf_for_angiography_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from nerf_for_angiography_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/angio_b200.h but not exported"
    # the ctypes prototype table covers the header exactly
    assert sorted(_lib.PROTOTYPES) == _declared_symbols()
    assert _lib.load().angio_version() == 100


def test_argument_validation_without_gpu():
    """error convention: negative library codes + a message, no exceptions across the ABI"""
    from nerf_for_angiography_b200 import _lib
    lib = _lib.load()
    bad = _lib.MlpDesc(1, 0, 128, 4)                       # fourier with basis 0
    assert lib.angio_mlp_param_count(ctypes.byref(bad)) == _lib.ERR_INVALID_ARG
    assert b"invalid" in lib.angio_last_error_string()
    good = _lib.MlpDesc(1, 5, 128, 4)
    assert lib.angio_mlp_param_count(ctypes.byref(good)) == 70548 - 4   # reference count minus the dead img1/img2 (2+2)
    assert lib.angio_mlp_input_width(ctypes.byref(good)) == 33
    assert lib.angio_mlp_param_count(ctypes.byref(_lib.MlpDesc(0, 0, 128, 4))) == 66693 - 4
    assert lib.angio_mlp_param_count(ctypes.byref(_lib.MlpDesc(1, 5, 256, 8))) == 535316 - 4
    rc = lib.angio_raygen(None, 0, None, None, None, 10, 4, 4, 1.0, None, None, None, None, None)
    assert rc == _lib.ERR_INVALID_ARG
    with pytest.raises(RuntimeError):
        _lib.check(rc, "angio_raygen")
    # entry points added for the sync-free / multi-GPU path reject bad arguments the same way (no GPU work is attempted)
    assert lib.angio_sample_rays(None, 100, 200, 1, 1.0, 300, None, None, None, 0, None) == _lib.ERR_INVALID_ARG       # n > n_pool
    assert lib.angio_sample_rays_workspace_bytes(0, 10) == _lib.ERR_INVALID_ARG
    assert lib.angio_sample_rays_workspace_bytes(1000, 100) > 1000 * 12
    assert lib.angio_march_runs_bytes(65536) == (3 * 8 + 1) * 65536 * 4
    assert lib.angio_ray_segment_counts(None, 10, 0, 32, None, None, None) == _lib.ERR_INVALID_ARG
    assert lib.angio_visibility_head(None, None, 10, 0, 0.01, None, None) == _lib.ERR_INVALID_ARG
    assert lib.angio_project_volume(None, 8, 8, 8, None, None, None, 10, None, 4, 1, None, None) == _lib.ERR_INVALID_ARG
    assert lib.angio_signal_peers(None, 2, 0, 1, None) == _lib.ERR_INVALID_ARG
    assert lib.angio_adam_step_allreduce(None, None, 2, None, 1, None, None, 10, 1e-4, 0.9, 0.999, 1e-8, 1, 1.0, -1, None, None) == _lib.ERR_INVALID_ARG
    assert b"angio_adam_step_allreduce" in lib.angio_last_error_string()
    # lazy marching entry points
    assert lib.angio_march_head(None, None, 10, None, None, 32, None, 0.0, 1.0, 0.1, 32, None, None, None, None, None, None, None, None,
                                None) == _lib.ERR_INVALID_ARG
    assert lib.angio_visibility_head_mask(None, None, None, 10, 32, 0.01, 0.0, None, None, None, None, None, None) == _lib.ERR_INVALID_ARG
    assert lib.angio_compact_head_tail(None, None, None, None, None, None, None, None, None, None, 10, 0, None, None, None, None) == _lib.ERR_INVALID_ARG


def test_python_surface_fails_loudly_on_cpu_tensors():
    import torch
    from nerf_for_angiography_b200 import ops
    with pytest.raises(ValueError):
        ops.composite_forward(torch.zeros(4), torch.zeros(4), torch.ones(4), torch.tensor([0, 4], dtype=torch.int32))


def test_buffer_pool_is_grow_only_and_stable():
    """Host logic of the scratch pool behind the sync-free loop: a request that fits returns the SAME storage (no allocation in
    steady state), growth doubles, reserve() is exact, typed() carves dtype views."""
    import torch
    from nerf_for_angiography_b200.ops import BufferPool
    cpu = torch.device("cpu")
    pool = BufferPool()
    a = pool.get("x", 1000, cpu)
    assert a.dtype == torch.uint8 and a.numel() >= 2000
    assert pool.get("x", 1500, cpu).data_ptr() == a.data_ptr()            # fits: same block
    b = pool.get("x", 5000, cpu)
    assert b.numel() >= 10000                                             # grew by doubling the request
    assert pool.get("x", 10, cpu).data_ptr() == b.data_ptr()              # never shrinks
    r = pool.reserve("saved", 4096, cpu)
    assert 4096 <= r.numel() <= 4096 + 256 and pool.reserve("saved", 100, cpu).data_ptr() == r.data_ptr()
    t = pool.typed("ids", 100, torch.int32, cpu)
    assert t.dtype == torch.int32 and t.numel() == 100
    t[:] = 7
    assert pool.typed("ids", 50, torch.int32, cpu).data_ptr() == t.data_ptr() and int(pool.typed("ids", 50, torch.int32, cpu)[49]) == 7
    assert pool.typed("f", 3, torch.float32, cpu).dtype == torch.float32
