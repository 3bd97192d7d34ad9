"""Free functions of the reference's `nerf` namespace that sit on the hot path (src/nerf/__init__.py:1-5)."""
from __future__ import annotations

import numpy as np
import torch

from .engine import Engine, RenderSettings

_util_engine = None


def _engine() -> Engine:
    """A weight-less handle for the stand-alone ray-generation entry points."""
    global _util_engine
    if _util_engine is None:
        _util_engine = Engine(dict(num_layers=2, hidden_size=128), None, RenderSettings(num_coarse=8, num_fine=0))
    return _util_engine


def meshgrid_xy(tensor1, tensor2):
    """src/nerf/nerf_helpers.py:184-196 — index bookkeeping only (no arithmetic), kept in torch."""
    ii, jj = torch.meshgrid(tensor1, tensor2, indexing="ij")
    return ii.transpose(-1, -2), jj.transpose(-1, -2)


def get_ray_bundle(height: int, width: int, focal_length: float, tform_cam2world: torch.Tensor):
    """src/nerf/nerf_helpers.py:226-277 on the device: returns (ray_origins (3,), ray_directions (H,W,3))."""
    return _engine().ray_bundle(tform_cam2world, height, width, float(focal_length))


def ndc_rays(H, W, focal, near, rays_o=None, rays_d=None, tform_cam2world=None):
    """src/nerf/nerf_helpers.py:280-307.  Called like the reference — ndc_rays(H, W, focal, near, rays_o, rays_d), the
    positional form of DataBundle.ndc (src/data/data_helpers.py:164-167) — the given rays are warped on the device and come
    back on the device of `rays_d` (CPU tensors in, CPU tensors out).  With `tform_cam2world=` instead of rays, the pinhole
    rays are generated and warped in one kernel (the fused fast path nm_render_image uses)."""
    if rays_o is not None and rays_d is not None:
        rays_d = torch.as_tensor(rays_d)
        o, d = _engine().ndc_rays(H, W, float(focal), float(near), torch.as_tensor(rays_o), rays_d)
        return (o, d) if rays_d.is_cuda else (o.cpu(), d.cpu())
    if tform_cam2world is None:
        raise ValueError("ndc_rays needs (rays_o, rays_d) or tform_cam2world=")
    return _engine().ray_bundle(tform_cam2world, H, W, float(focal), ndc=True, ndc_near=float(near))


def pose_spherical(theta, phi, radius):
    """src/data/data_helpers.py:10-37 (host-side, 16 numbers; used to synthesise benchmark poses)."""
    def trans(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], dtype=np.float32)

    def rot_phi(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], dtype=np.float32)

    def rot_theta(th):
        return np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=np.float32)

    c2w = rot_theta(theta / 180.0 * np.pi) @ (rot_phi(phi / 180.0 * np.pi) @ trans(radius))
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]) @ c2w
    return c2w.astype(np.float32)
