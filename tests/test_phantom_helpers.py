"""CPU-only: host-side phantom helpers against golden vectors from the reference's own
phantomdata/helpers.py (tests/golden/make_golden.py)."""
import os

import numpy as np

from nerf_for_angiography_b200.data import get_weighted_img, transfer_func_ct


def test_ct_transfer_function_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "phantom.npz"))
    assert np.array_equal(transfer_func_ct(g["hu"]), g["tf"])                       # same float64 arithmetic: bit-exact
    assert np.array_equal(transfer_func_ct(g["hu"], binary=True), g["tf_bin"])
    # known answers at the knots (helpers.py:36-58)
    assert np.allclose(transfer_func_ct([-5.0, 0.0, 1585.85, 2332.9, 3306.18, 4000.0, 9000.0]), [0, 0, 0.05, 0, 0.2, 0.4, 0.4], atol=1e-12)
    assert transfer_func_ct(np.array([3000], dtype=np.int16)).dtype == np.float64    # integer HU volumes are promoted like the reference does


def test_weight_image_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "phantom.npz"))
    w = get_weighted_img(g["img"])
    assert np.array_equal(w, g["wimg"])
    assert w.min() == 1e-10 and np.isclose(w.max(), 1.0) and w.shape == g["img"].shape
    assert np.all(w[g["img"] >= 1] == 1e-10)                                          # background: distance 0


def test_packed_ray_indices_keeps_its_layout_through_copies_only():
    """nerfacc.ray_marching returns int64 ray_indices that carry the packed segment offsets for the compositor
    (nerf_helpers_acc.acc_render_volume_density): same-sample copies keep them, anything that changes the sample set drops
    them (a stale layout would composite the wrong segments) and never leaks the subclass into derived tensors."""
    import copy
    import io
    import torch
    from nerf_for_angiography_b200.nerfacc import PackedRayIndices
    from nerf_for_angiography_b200.nerf.nerf_helpers_acc import packed_offsets
    idx32 = torch.tensor([0, 0, 1, 3, 3, 3], dtype=torch.int32)
    off = torch.tensor([0, 2, 3, 3, 6], dtype=torch.int32)
    r = PackedRayIndices.wrap(idx32.long(), off, idx32)
    for f in (lambda t: t.clone(), lambda t: t.detach(), lambda t: t.contiguous(), lambda t: t.long(), lambda t: t.to(torch.int64),
              lambda t: t.to("cpu"), copy.deepcopy):
        o = f(r)
        assert isinstance(o, PackedRayIndices) and o._angio_offsets.tolist() == off.tolist() and o.tolist() == r.tolist()
        assert packed_offsets(o, 4) is o._angio_offsets
    for f in (lambda t: t[1:], lambda t: t[t > 0], lambda t: t + 1, lambda t: t.int(), lambda t: torch.arange(10.0)[t], lambda t: t.reshape(-1, 1),
              lambda t: t.float(), lambda t: t[1:] >= t[:-1]):
        o = f(r)
        assert type(o) is torch.Tensor and not hasattr(o, "_angio_offsets")
    assert packed_offsets(r[2:], 4).tolist() == [0, 0, 1, 1, 4]            # a slice falls back to the search
    buf = io.BytesIO()
    torch.save(r, buf)
    buf.seek(0)
    back = torch.load(buf)
    assert type(back) is torch.Tensor and back.tolist() == r.tolist()
