"""Host-side glue of src/nerf/nerf_helpers.py that eval_nerf.py / mesh_nerf.py import by name."""
import numpy as np
import torch

from nerfmeshes_b200.mesh import export_obj  # noqa: F401  (byte-identical OBJ text, tests/golden/golden_mesh.obj)
from nerfmeshes_b200.nerf_api import get_ray_bundle, meshgrid_xy, ndc_rays  # noqa: F401

try:
    from tqdm import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None


def mse2psnr(mse):
    """PSNR for a peak of 1.0; a zero loss is floored to 1e-5 (src/nerf/nerf_helpers.py:17-23)."""
    mse = torch.as_tensor(mse, dtype=torch.float32)
    if float(mse) == 0.0:
        mse = torch.tensor(1e-5)
    return -10.0 * torch.log10(mse)


def batchify(*data, batch_size=1024, device="cpu", progress=True):
    """Yield aligned slices of the given tensors moved to `device` (src/nerf/nerf_helpers.py:114-139)."""
    n = data[0].shape[0]
    assert all(t is None or t.shape[0] == n for t in data), "Sizes of tensors must match for dimension 0."

    def gen():
        for s in range(0, n, batch_size):
            yield [None if t is None else t[s:s + batch_size].to(device) for t in data]

    total = (n - 1) // batch_size + 1
    return _tqdm(gen(), total=total) if (progress and _tqdm is not None) else gen()


def cast_to_pil_image(tensor):
    """(H,W,3) float in [0,1] -> uint8 (H,W,3) array, rounding like torchvision's ToPILImage (x*255 -> byte)."""
    t = tensor.detach().cpu().float()
    return t.mul(255).byte().numpy()


def cast_to_image(tensor):
    return np.moveaxis(cast_to_pil_image(tensor), [-1], [0])


def cast_to_disparity_image(tensor, white_background=False):
    img = (tensor - tensor.min()) / (tensor.max() - tensor.min())
    img = (img.clamp(0.0, 1.0) * 255).byte()
    if white_background:
        img[img == 0] = 255
    return img.detach().cpu().numpy()


def export_point_cloud(it, ray_origins, ray_directions, depth_fine, dep_target):
    """Two-colour point cloud OBJ of predicted (red) vs target (blue) depths (src/nerf/nerf_helpers.py:141-152)."""
    def pts(depth):
        return (ray_origins + ray_directions * depth[..., None]).view(-1, 3)

    out, tgt = pts(depth_fine), pts(dep_target)
    col = torch.zeros(out.shape[0] + tgt.shape[0], 3)
    col[: out.shape[0], 0] = 1.0
    col[out.shape[0]:, 2] = 1.0
    nrm = -ray_directions.reshape(-1, 3)
    export_obj(torch.cat((out, tgt)), [], col, torch.cat((nrm, nrm)), f"{it:04d}.obj")
