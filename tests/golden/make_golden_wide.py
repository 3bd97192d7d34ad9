"""Wide end-to-end goldens: 4096 rays x 3 poses per shipped checkpoint through the UNMODIFIED reference.

    python tests/golden/make_golden_wide.py            # ~2 min on 8 CPU threads

The stage-level fixtures of make_golden.py cover 64-96 rays of one pose; these cover 12,288 rays per checkpoint (half of them
through the object, half anywhere in the image, three poses) with only the per-ray outputs kept (0.5 MB per checkpoint):
rgb / depth / acc / disp of the final bundle, the coarse rgb, and depth_raw = sum(w * t) before the eval-mode threshold
(depth[acc < 1] = 0 flips on 1-ulp changes).  The pipeline is run with the reference's own modules, call for
call like NeRFModel.forward (src/models/model_nerf.py:52-76) / BuFFModel.forward (src/models/model_buff.py:34-69), so that
t_fine is available for depth_raw; the first 256 rays of every pose are also pushed through model.forward itself and must
agree bit for bit (asserted here).  Also writes golden_mesh_inputs.npz / golden_mesh.obj (the reference's export_obj on a
small synthetic mesh: tests of the native OBJ writer)."""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

rh.install()
import nerf  # noqa: E402
import models  # noqa: E402
from data.data_helpers import pose_spherical  # noqa: E402

torch.set_num_threads(8)
META = dict(torch_version=torch.__version__, threads=torch.get_num_threads())
N_RAYS = 4096


def pick_rays(H, W, n, seed):
    g = torch.Generator().manual_seed(seed)
    r1 = torch.randint(H // 4, 3 * H // 4, (n // 2,), generator=g)
    c1 = torch.randint(W // 4, 3 * W // 4, (n // 2,), generator=g)
    r2 = torch.randint(0, H, (n - n // 2,), generator=g)
    c2 = torch.randint(0, W, (n - n // 2,), generator=g)
    return torch.cat([r1 * W + c1, r2 * W + c2])


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()},
                        meta=np.array(str(META)))
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


@torch.no_grad()
def nerf_pipeline(m, o, dirs, bounds, chunk=1024):
    out = {k: [] for k in ("rgb", "depth", "depth_raw", "acc", "disp", "coarse_rgb")}
    near, far = bounds
    for s in range(0, dirs.shape[0], chunk):
        d = dirs[s:s + chunk]
        oo = o[s:s + chunk] if o.dim() == 2 and o.shape[0] == dirs.shape[0] else o
        cfgv = m.cfg.nerf.validation
        t_c = m.sampler(cfgv, d.shape[0], near, far)
        p_c = models.intervals_to_ray_points(t_c, d, oo)
        cb = m.volume_renderer(m.model_coarse(p_c, d[..., None, :].expand_as(p_c)), t_c, d)
        t_f = m.sample_pdf(t_c, cb.weights, cfgv.perturb)
        p_f = models.intervals_to_ray_points(t_f, d, oo)
        fb = m.volume_renderer(m.model_fine(p_f, d[..., None, :].expand_as(p_f)), t_f, d)
        if s == 0:
            c2, f2 = m.forward((oo[:256] if oo.dim() == 2 and oo.shape[0] == d.shape[0] else oo, d[:256], bounds))
            assert torch.equal(f2.rgb_map, fb.rgb_map[:256]) and torch.equal(c2.rgb_map, cb.rgb_map[:256]), "manual pipeline != model.forward"
        for k, v in (("rgb", fb.rgb_map), ("depth", fb.depth_map), ("depth_raw", (fb.weights * t_f).sum(-1)), ("acc", fb.acc_map),
                     ("disp", fb.disp_map), ("coarse_rgb", cb.rgb_map)):
            out[k].append(v)
    return {k: torch.cat(v) for k, v in out.items()}


@torch.no_grad()
def buff_pipeline(m, o, dirs, bounds, chunk=1024):
    out = {k: [] for k in ("rgb", "depth", "depth_raw", "acc", "disp", "ray_mask")}
    near, far = bounds
    for s in range(0, dirs.shape[0], chunk):
        d = dirs[s:s + chunk]
        z, _, mask = m.tree.batch_ray_voxel_intersect(o[None], d, near, far, samples_count=192)
        t_u = m.sampler(m.cfg.nerf.validation, d.shape[0], near, far)
        z[~mask] = t_u[~mask]
        p = models.intervals_to_ray_points(z, d, o[None])
        b = m.volume_renderer(m.model(p, d[..., None, :].expand_as(p)), z, d)
        if s == 0:
            b2 = m.forward((o[None], d[:256], bounds))
            assert torch.equal(b2.rgb_map, b.rgb_map[:256]), "manual pipeline != model.forward"
        for k, v in (("rgb", b.rgb_map), ("depth", b.depth_map), ("depth_raw", (b.weights * z).sum(-1)), ("acc", b.acc_map),
                     ("disp", b.disp_map), ("ray_mask", mask)):
            out[k].append(v)
    return {k: torch.cat(v) for k, v in out.items()}


@torch.no_grad()
def main():
    H = W = 800
    focal = float(0.5 * 800 / np.tan(0.5 * 0.6911112))
    angles = np.linspace(-270, 90, 120, endpoint=False)[[0, 40, 85]]
    poses = [torch.from_numpy(pose_spherical(float(a), -30.0, 4.0)) for a in angles]
    bounds = torch.tensor([2.0, 6.0])

    m = rh.load_model("NeRFModel", "colab-lego-nerf-high-res")
    res = []
    for i, pose in enumerate(poses):
        o, d = nerf.get_ray_bundle(H, W, focal, pose)
        ids = pick_rays(H, W, N_RAYS, 100 + i)
        r = nerf_pipeline(m, o, d.view(-1, 3)[ids].contiguous(), bounds)
        r["ray_ids"] = ids
        res.append(r)
    save("golden_wide_lego.npz", H=H, W=W, focal=focal, poses=torch.stack(poses), bounds=bounds,
         **{k: torch.stack([r[k] for r in res]) for k in res[0]})

    mb = rh.load_model("BuFFModel", "buff-synthetic-lego")
    res = []
    for i, pose in enumerate(poses):
        o, d = nerf.get_ray_bundle(H, W, focal, pose)
        ids = pick_rays(H, W, N_RAYS, 200 + i)
        r = buff_pipeline(mb, o, d.view(-1, 3)[ids].contiguous(), bounds)
        r["ray_ids"] = ids
        res.append(r)
    save("golden_wide_buff.npz", H=H, W=W, focal=focal, poses=torch.stack(poses), bounds=bounds,
         **{k: torch.stack([r[k] for r in res]) for k in res[0]})

    mf = rh.load_model("NeRFModel", "nerf-colmap-fern")
    Hf, Wf, ff = 756, 1008, 815.13
    fposes = []
    for dx, dy in ((0.0, 0.0), (0.1, 0.0), (-0.1, 0.1)):
        p = torch.eye(4)
        p[0, 3], p[1, 3] = dx, dy
        fposes.append(p)
    bf = torch.tensor([0.0, 1.0])
    res = []
    for i, pose in enumerate(fposes):
        o, d = nerf.get_ray_bundle(Hf, Wf, ff, pose)
        on, dn = nerf.ndc_rays(Hf, Wf, ff, 1.0, o[None, None, :].expand(Hf, Wf, 3), d)
        ids = pick_rays(Hf, Wf, N_RAYS, 300 + i)
        r = nerf_pipeline(mf, on.reshape(-1, 3)[ids].contiguous(), dn.reshape(-1, 3)[ids].contiguous(), bf)
        r["ray_ids"] = ids
        res.append(r)
    save("golden_wide_fern.npz", H=Hf, W=Wf, focal=ff, poses=torch.stack(fposes), bounds=bf,
         **{k: torch.stack([r[k] for r in res]) for k in res[0]})

    # ---------------------------------------------------------------- the reference's OBJ writer on a small synthetic mesh
    g = torch.Generator().manual_seed(5)
    verts = (torch.rand(7, 3, generator=g) * 2.4 - 1.2).float()
    verts[0, 0], verts[1, 1], verts[2, 2] = 1e-5, -3.0e-7, 123456.0
    normals = torch.nn.functional.normalize(torch.randn(7, 3, generator=g), dim=-1)
    diffuse = torch.rand(7, 3, generator=g).numpy().astype(np.float32)
    tris = torch.tensor([[0, 1, 2], [2, 3, 4], [4, 5, 6], [6, 0, 3]], dtype=torch.int32)
    cwd = os.getcwd()
    os.chdir(HERE)
    try:
        nerf.export_obj(verts, tris, diffuse, normals, "golden_mesh.obj")
    finally:
        os.chdir(cwd)
    np.savez(os.path.join(HERE, "golden_mesh_inputs.npz"), v=verts.numpy(), n=normals.numpy(), d=diffuse, f=tris.numpy())
    print("golden_mesh.obj written by the reference's export_obj")


if __name__ == "__main__":
    main()
