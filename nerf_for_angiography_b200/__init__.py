"""nerf_for_angiography_b200 -- B200-native (sm_100a) training / rendering hot path of
kirstenmaas/nerf-for-angiography behind the reference's own call surface.

    from nerf_for_angiography_b200 import CPPN, OccupancyGrid, ContractionType
    from nerf_for_angiography_b200 import acc_ray_marching, acc_render_volume_density, acc_update_n_step
    from nerf_for_angiography_b200 import get_predictions, sample_pixel_rays, render_rays

Everything below these names runs in hand-written CUDA kernels reached through the C ABI of
``libangio_b200.so`` (``include/angio_b200.h``); there is no CPU or eager-PyTorch fallback.
"""
from . import _lib, ops  # noqa: F401
from .model.CPPN import CPPN  # noqa: F401
from .nerfacc import ContractionType, OccupancyGrid, ray_marching  # noqa: F401
from .nerf.nerf_helpers_acc import acc_ray_marching, acc_render_volume_density, acc_update_n_step  # noqa: F401
from .nerf.nerf_helpers import get_predictions, sample_pixel_rays  # noqa: F401
from .render import render_rays  # noqa: F401
from .geometry import get_ray_values, source_matrix  # noqa: F401
from .inference import query_volume, render_projections  # noqa: F401

__version__ = "0.1.0"
