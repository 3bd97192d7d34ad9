"""End-to-end training on the fused path, without a dataset: a fresh two-network NeRF (8x256) is fitted to images rendered
by the shipped lego checkpoint (re-packed in tests/golden/weights_lego_nerf.npz) — the reference's training loop
(training_step -> Adam -> exponential LambdaLR, src/models/model_base.py:150-177) with every forward / backward on the
library's kernels.  Prints the loss / PSNR curve and the throughput.      python tools/train_demo.py [steps] [rays] [seed]
As with the reference, an unlucky initialisation can leave one of the two networks in NeRF's "all-empty" local minimum for
a while (the reference carries check_early_stopping for it, src/models/model_base.py:179-186); the gradients themselves are
checked against torch autograd in tests/test_gpu_train.py and tools/train_diag.py.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import nerfmeshes_b200 as nm

ARCH = dict(num_layers=8, hidden_size=256, skip_step=4, num_encoding_fn_xyz=10, num_encoding_fn_dir=4, include_input_xyz=True,
            include_input_dir=True, log_sampling_xyz=True, log_sampling_dir=True, use_viewdirs=True)
CFG = {"dataset.near": 2.0, "dataset.far": 6.0, "dataset.white_background": False,
       "models.coarse_type": "FlexibleNeRFModel", "models.fine_type": "FlexibleNeRFModel", "models.use_fine": True,
       **{f"models.coarse.{k}": v for k, v in ARCH.items()}, **{f"models.fine.{k}": v for k, v in ARCH.items()}}
for mode in ("train", "validation"):
    CFG.update({f"nerf.{mode}.num_coarse": 64, f"nerf.{mode}.num_fine": 128, f"nerf.{mode}.perturb": mode == "train",
                f"nerf.{mode}.lindisp": False, f"nerf.{mode}.radiance_field_noise_std": 0.2 if mode == "train" else 0.0})


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    raw = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_lego_nerf.npz"))
    z = {k: torch.from_numpy(raw[k]) for k in raw.files if raw[k].dtype.kind in "fiub"}
    teacher = nm.NeRFModel.from_npz(CFG, z).cuda().eval()
    torch.manual_seed(int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    student = nm.NeRFModel(CFG).cuda().train()
    opt = torch.optim.Adam(student.parameters(), lr=5e-4)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda s: 0.1 ** (s / 250000))
    H = W = 200
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    poses = [nm.pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, 24, endpoint=False)]
    views = []
    with torch.no_grad():
        for p in poses:                                   # "dataset": rays + teacher colours, resident on the device
            o, d = student._engine().ray_bundle(p, H, W, focal)
            rgb = teacher._engine().render_image(p, H, W, focal, 2.0, 6.0, want=["rgb"])["rgb"]
            views.append((o, d.reshape(-1, 3), rgb))
    g = torch.Generator(device="cuda").manual_seed(1)
    torch.cuda.synchronize()
    t0 = time.time()
    for step in range(steps):
        o, d, rgb = views[step % len(views)]
        sel = torch.randint(0, d.shape[0], (R,), device="cuda", generator=g)
        opt.zero_grad(set_to_none=True)
        out = nm.training_step(student, (o, d[sel], (2.0, 6.0)), rgb[sel], global_step=step)
        opt.step()
        sched.step()
        if step % max(steps // 10, 1) == 0 or step == steps - 1:
            print(f"step {step:5d}  loss {out['loss']:.5f}  coarse psnr {out['log']['train/coarse_psnr']:.2f}  "
                  f"fine psnr {out['log']['train/fine_psnr']:.2f}", flush=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    student.eval()
    with torch.no_grad():
        p = nm.pose_spherical(15.0, -30.0, 4.0)
        a = student._engine().render_image(p, H, W, focal, 2.0, 6.0, want=["rgb"])["rgb"]
        b = teacher._engine().render_image(p, H, W, focal, 2.0, 6.0, want=["rgb"])["rgb"]
        mse = float(torch.mean((a - b) ** 2))
    print(f"{steps} steps x {R} rays in {dt:.1f} s = {steps * R / dt:,.0f} rays/s wall (forward + backward + Adam + weight re-pack); "
          f"held-out view PSNR vs the teacher {-10 * np.log10(mse):.2f} dB")


if __name__ == "__main__":
    main()
