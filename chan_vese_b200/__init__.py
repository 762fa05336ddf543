"""chan_vese_b200 -- B200-native (sm_100a) Perona-Malik + Chan-Sandberg-Vese solvers behind a C ABI.

Layout: csrc/ (CUDA kernels + the C ABI of include/chan_vese_b200.h), build.py (in-tree nvcc build),
_ffi.py (ctypes declaration of the ABI), solver.py (host-side mirror of the reference's solver interface),
frontend.py (the reference's command line with cv2 image I/O), synth.py (deterministic synthetic inputs).
"""
from .solver import (Batch, ChanVeseError, Context, ParallelPixelFunction, Region, Session, auto_tile_rows, chan_vese,
                     curvature, default_context, levelset_checkerboard, levelset_circ, levelset_rect, make_params,
                     perona_malik, pm_num_steps, region_variance, separate, separate_mask, slab_partition)

__all__ = ["Batch", "ChanVeseError", "Context", "ParallelPixelFunction", "Region", "Session", "auto_tile_rows",
           "chan_vese", "curvature", "default_context", "levelset_checkerboard", "levelset_circ", "levelset_rect",
           "make_params", "perona_malik", "pm_num_steps", "region_variance", "separate", "separate_mask",
           "slab_partition"]
