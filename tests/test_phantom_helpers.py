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
