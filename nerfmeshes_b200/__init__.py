"""nerfmeshes_b200 — B200-native (sm_100a) NeRF render / dense-grid hot path of qway/nerfmeshes.

Layout: csrc/ (CUDA kernels + C ABI, built into lib/libnerfmeshes_b200.so), _lib.py (ctypes binding), engine.py (handle +
torch plumbing), models.py / nerf_api.py / mesh.py / train.py / eval.py (host mirror of the reference's interface for this path).
"""
from . import _lib
from ._lib import NmError, PREC_EXACT, PREC_FAST, PREC_FP32
from .cfgnode import CfgNode, flatten_dict, nest_dict
from .engine import Engine, RenderSettings
from .models import (BaseModel, BuFFModel, FlexibleNeRFModel, NeRFModel, OutputBundle, TreeSampling,
                     load_lightning_checkpoint)
from .nerf_api import get_ray_bundle, meshgrid_xy, ndc_rays, pose_spherical
from .mesh import extract_geometry, extract_iso_level, extract_radiance, marching_cubes
from .train import training_step

__all__ = [n for n in dir() if not n.startswith("_")]
