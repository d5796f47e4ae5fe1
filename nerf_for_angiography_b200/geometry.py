"""C-arm cone-beam geometry: the reference's camera model (/root/reference/phantomdata/proj_helpers.py:34-77) and
pixel -> ray mapping (/root/reference/phantomdata/helpers.py:156-175).

The 4x4 matrices are per-view host work in float64 (numpy), exactly as in the reference; the per-pixel work
(W*H rays per view) runs in the ``angio_raygen`` kernel, which evaluates the reference's float64 expression and
rounds once to float32 -- bit-identical to ``get_ray_values(...)[k].float()``.
"""
import numpy as np
import torch

from . import ops


def x_rotation_matrix(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def y_rotation_matrix(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def z_rotation_matrix(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)


def translation_matrix(vec):
    m = np.identity(4)
    m[:3, 3] = np.asarray(vec, dtype=np.float64)[:3]
    return m


def get_rotation(theta, phi, larm, type='rotation'):
    return np.linalg.inv(z_rotation_matrix(np.deg2rad(larm)).dot(
        x_rotation_matrix(np.deg2rad(theta)).dot(y_rotation_matrix(np.deg2rad(phi)))))


def source_matrix(source_pt, theta, phi, larm=0, translation=[0, 0, 0], type='rotation'):
    m2 = get_rotation(theta, phi, larm)
    m3 = translation_matrix(source_pt)
    m4 = translation_matrix([translation[0], translation[1], translation[2], 1])
    return m4.dot(m2.dot(m3))


def get_ray_values(theta, phi, larm, src_pt, img_width, img_height, focal_length, device, translation=np.array([0, 0, 0])):
    """Reference signature.  Returns (ray_origins[H,W,3], ray_directions[H,W,3], src_matrix, ii, jj) with the rays
    already in float32 (the reference casts with .float() before use)."""
    src_matrix = source_matrix(src_pt, theta, phi, larm, translation)
    cam = torch.from_numpy(src_matrix[None].copy()).to(device)
    o, d = ops.raygen(cam, int(img_width), int(img_height), float(focal_length), view=0)
    ii, jj = torch.meshgrid(torch.arange(0, img_width, device=device, dtype=torch.float64),
                            torch.arange(0, img_height, device=device, dtype=torch.float64), indexing='xy')
    return o.view(int(img_height), int(img_width), 3), d.view(int(img_height), int(img_width), 3), src_matrix, ii, jj
