"""Device-resident evaluation loop — the body of eval_nerf.eval_nerf (src/eval_nerf.py:50-105) with the per-chunk
H2D/D2H traffic removed: one fused call per image (pose in, maps out), images/disparities written like the reference."""
from __future__ import annotations

import os
from typing import Iterable, Optional

import numpy as np
import torch


def cast_to_pil_image(t: torch.Tensor) -> np.ndarray:
    """(H,W,3) float in [0,1] -> uint8 array (src/nerf/nerf_helpers.py:155-160: ToPILImage == x*255 -> byte)."""
    return t.detach().cpu().float().mul(255).byte().numpy()


def cast_to_disparity_image(t: torch.Tensor, white_background=False) -> np.ndarray:
    """src/nerf/nerf_helpers.py:171-181."""
    img = (t - t.min()) / (t.max() - t.min())
    img = (img.clamp(0.0, 1.0) * 255).byte()
    if white_background:
        img[img == 0] = 255
    return img.detach().cpu().numpy()


def mse2psnr(mse: torch.Tensor) -> torch.Tensor:
    mse = torch.as_tensor(mse, dtype=torch.float32)
    return -10.0 * torch.log10(torch.where(mse == 0, torch.full_like(mse, 1e-5), mse))


def eval_poses(model, poses: Iterable, H: int, W: int, focal: float, near: float, far: float, *, ndc=False,
               targets: Optional[Iterable[torch.Tensor]] = None, save_dir: Optional[str] = None, save_disparity=False):
    """Render every pose; returns dict(rgb=[(H,W,3) cpu tensors], disp=[(H,W)], mse=[...], psnr=[...]).
    `targets`: optional iterable of (H,W,3) images -> per-image MSE/PSNR like eval_nerf.py:73-76,101-105."""
    eng = model._engine()
    out = dict(rgb=[], disp=[], mse=[], psnr=[])
    targets = list(targets) if targets is not None else None
    if save_dir:
        os.makedirs(os.path.join(save_dir, "images"), exist_ok=True)
        if save_disparity:
            os.makedirs(os.path.join(save_dir, "disparity"), exist_ok=True)
    for i, pose in enumerate(poses):
        o = eng.render_image(pose, H, W, focal, near, far, ndc=ndc, want=["rgb", "disp"])
        rgb, disp = o["rgb"].view(H, W, 3), o["disp"].view(H, W)
        if targets is not None:
            mse = torch.nn.functional.mse_loss(rgb, targets[i].to(rgb.device).view(H, W, 3))
            out["mse"].append(float(mse))
            out["psnr"].append(float(mse2psnr(mse)))
        out["rgb"].append(rgb.cpu())
        out["disp"].append(disp.cpu())
        if save_dir:
            from PIL import Image
            Image.fromarray(cast_to_pil_image(rgb)).save(os.path.join(save_dir, "images", f"{i:04d}.png"))
            if save_disparity:
                Image.fromarray(cast_to_disparity_image(disp, white_background=True)).save(
                    os.path.join(save_dir, "disparity", f"{i:04d}.png"))
    return out
