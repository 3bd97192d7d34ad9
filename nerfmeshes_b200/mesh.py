"""Dense-grid density query + iso-surface extraction — mirror of src/mesh_nerf.py:27-92 on the fused path."""
from __future__ import annotations

import numpy as np
import torch


def extract_radiance(model, args, device, nums, sigma_only=False, slab=None):
    """src/mesh_nerf.py:27-53.  Returns the (n0,n1,n2,4) numpy radiance grid like the reference, or — with
    sigma_only — the (n0,n1,n2) raw-density DEVICE tensor (the fast path extract_geometry uses; the reference
    discards the rgb channels anyway, mesh_nerf.py:73)."""
    assert isinstance(nums, (tuple, list, int)), "Nums arg should be either iterable or int."
    if isinstance(nums, int):
        nums = (nums,) * 3
    else:
        assert len(nums) == 3, "Nums arg should be of length 3, number of axes for 3D"
    tiles = [torch.linspace(-args.limit, args.limit, num) for num in nums]
    eng = model._engine()
    x0, x1 = slab if slab is not None else (0, nums[0])
    if sigma_only:
        return eng.grid_sigma(tiles, x0, x1)
    sigma, rgb = eng.grid_sigma(tiles, x0, x1, with_rgb=True)
    return torch.cat((rgb, sigma[..., None]), -1).cpu().numpy()


def extract_iso_level(density, args, engine=None):
    """src/mesh_nerf.py:56-65: clamp(iso_level, min+std, max-std)."""
    if isinstance(density, torch.Tensor) and density.is_cuda:
        mn, mx, sd = engine.volume_stats(density)
        mn, mx, sd = np.float32(mn), np.float32(mx), np.float32(sd)
    else:
        density = np.asarray(density)
        mn, mx, sd = density.min(), density.max(), density.std()
    return min(max(args.iso_level, mn + sd), mx - sd)


def marching_cubes(volume, level, engine=None):
    """skimage.measure.marching_cubes(volume, level) seam (src/mesh_nerf.py:79): (verts, faces, normals, values)."""
    if engine is None:
        from .nerf_api import _engine
        engine = _engine()
    vol = torch.as_tensor(volume)
    verts, faces, normals = engine.marching_cubes(vol, float(level))
    values = torch.zeros(verts.shape[0])
    return verts.cpu().numpy(), faces.cpu().numpy(), normals.cpu().numpy(), values.numpy()


def extract_geometry(model, device, args):
    """src/mesh_nerf.py:68-92: sigma sweep -> adaptive iso -> marching cubes -> rescale to (-limit, limit)."""
    eng = model._engine()
    density = extract_radiance(model, args, device, args.res, sigma_only=True)
    iso_value = extract_iso_level(density, args, eng)
    verts, faces, normals = eng.marching_cubes(density, float(iso_value))
    # the reference rescales CPU tensors (:82-90); do the same on the host so the rounding is identical (torch's CUDA
    # division by a python scalar multiplies by the reciprocal, which differs in the last bit)
    vertices = args.limit * (verts.cpu() / (args.res / 2.0) - 1.0)    # keeps the reference's res/2 scale (:90)
    return vertices, faces.cpu(), normals.cpu(), density.cpu().numpy()
