"""Diagnostic: fused training_step vs the autograd route on the lego-sized nets at 4096 rays (same init, rays, seeds)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nerfmeshes_b200 as nm
from train_demo import CFG

R = 4096
torch.manual_seed(0)
a = nm.NeRFModel(CFG).cuda().train()
b = nm.NeRFModel(CFG).cuda().train()
b.load_state_dict(a.state_dict())
raw = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_lego_nerf.npz"))
z = {k: torch.from_numpy(raw[k]) for k in raw.files if raw[k].dtype.kind in "fiub"}
teacher = nm.NeRFModel.from_npz(CFG, z).cuda().eval()
H = W = 200
focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
p = nm.pose_spherical(30.0, -30.0, 4.0)
with torch.no_grad():
    o, d = a._engine().ray_bundle(p, H, W, focal)
    d = d.reshape(-1, 3)
    rgb = teacher._engine().render_image(p, H, W, focal, 2.0, 6.0, want=["rgb"])["rgb"]
oa = torch.optim.Adam(a.parameters(), lr=5e-4)
ob = torch.optim.Adam(b.parameters(), lr=5e-4)
g = torch.Generator(device="cuda").manual_seed(1)
for step in range(61):
    sel = torch.randint(0, d.shape[0], (R,), device="cuda", generator=g)
    oa.zero_grad(set_to_none=True); ob.zero_grad(set_to_none=True)
    out = nm.training_step(a, (o, d[sel], (2.0, 6.0)), rgb[sel], seed=1000 + step)
    c, f = b.forward((o, d[sel], (2.0, 6.0)), seed=1000 + step)
    lc = torch.nn.functional.mse_loss(c.rgb_map, rgb[sel]); lf = torch.nn.functional.mse_loss(f.rgb_map, rgb[sel])
    (lc + lf).backward()
    if step % 10 == 0:
        ga = a.model_coarse.layer1.weight.grad; gb = b.model_coarse.layer1.weight.grad
        fa = a.model_fine.layer1.weight.grad; fb = b.model_fine.layer1.weight.grad
        print(f"step {step:3d} fused: coarse {out['log']['train/coarse_loss']:.5f} fine {out['log']['train/fine_loss']:.5f} | autograd: coarse {lc.item():.5f} fine {lf.item():.5f} | "
              f"coarse grad norm {ga.norm().item():.3e} / {gb.norm().item():.3e} rel diff {((ga - gb).norm() / gb.norm()).item():.2e} | fine grad rel diff {((fa - fb).norm() / fb.norm()).item():.2e} | "
              f"w delta coarse {(a.model_coarse.layer1.weight - b.model_coarse.layer1.weight).abs().max().item():.2e}", flush=True)
    oa.step(); ob.step()
