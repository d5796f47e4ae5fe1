"""oracle/geometry.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

float64 numpy restatement of the reference's C-arm camera model:
  * rotation / translation / source matrices : /root/reference/phantomdata/proj_helpers.py:34-77
  * pixel -> cone-beam ray                   : /root/reference/phantomdata/helpers.py:156-175
Pinned against the reference code itself by tests/golden/make_golden.py.
"""
import numpy as np


def x_rotation_matrix(a):  # proj_helpers.py:34-40
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def y_rotation_matrix(a):  # proj_helpers.py:42-48
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def z_rotation_matrix(a):  # proj_helpers.py:50-56
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)


def translation_matrix(vec):  # proj_helpers.py:58-61
    m = np.identity(4)
    m[:3, 3] = np.asarray(vec, dtype=np.float64)[:3]
    return m


def get_rotation(theta, phi, larm):  # proj_helpers.py:63-66
    rz = z_rotation_matrix(np.deg2rad(larm))
    rx = x_rotation_matrix(np.deg2rad(theta))
    ry = y_rotation_matrix(np.deg2rad(phi))
    return np.linalg.inv(rz.dot(rx.dot(ry)))


def source_matrix(source_pt, theta, phi, larm=0, translation=(0, 0, 0)):  # proj_helpers.py:68-77
    m2 = get_rotation(theta, phi, larm)
    m3 = translation_matrix(source_pt)
    m4 = translation_matrix(translation)
    return m4.dot(m2.dot(m3))


def get_ray_values(theta, phi, larm, src_pt, img_width, img_height, focal_length, translation=(0, 0, 0)):
    """helpers.py:156-175.  Returns (origins[H,W,3], directions[H,W,3], M[4,4]) in float64.

    ii is the column (x) index and jj the row (y) index ('xy' meshgrid => arrays are [H, W]).
    d_k = sum_j dir_j * M[k, j]  with the three products added left to right.
    """
    M = source_matrix(src_pt, theta, phi, larm, translation)
    ii, jj = np.meshgrid(np.arange(img_width, dtype=np.float64), np.arange(img_height, dtype=np.float64),
                         indexing="xy")
    d0 = (ii - img_width / 2) / focal_length
    d1 = -(jj - img_height / 2) / focal_length
    d2 = -np.ones_like(ii)
    dirs = np.stack([d0, d1, d2], axis=-1)                       # [H, W, 3]
    prod = dirs[..., None, :] * M[:3, :3]                         # [H, W, 3(k), 3(j)]
    ray_d = (prod[..., 0] + prod[..., 1]) + prod[..., 2]          # left-to-right sum over j
    ray_o = np.broadcast_to(M[:3, 3], ray_d.shape).copy()
    return ray_o, ray_d, M
