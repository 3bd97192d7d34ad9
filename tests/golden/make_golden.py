"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py

Reads /root/reference (read-only): its source tree via ref_harness.py and the four shipped
checkpoints.  Writes (a) the checkpoint tensors re-packed as plain fp32 .npz (`weights_*.npz`; the
weight ABI of SURVEY Appendix A.1 — these are DATA, the PL pickles cannot travel to the GPU box) and
(b) input/output vectors of every hot-path stage (`golden_*.npz`).  torch version / thread count are
recorded in each file.  The reference ships no tests of its own, so these vectors are the
parity pin for oracle/nerf_oracle.py.
"""
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
import nerf  # noqa: E402  (reference package)
import models  # noqa: E402
import mesh_nerf  # noqa: E402
from data.data_helpers import pose_spherical  # noqa: E402

torch.set_num_threads(8)
META = dict(torch_version=torch.__version__, threads=torch.get_num_threads())


def npy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: npy(v) for k, v in arrs.items()},
                        meta=np.array(str(META)))
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def net_state(model, prefix):
    return {k[len(prefix):]: v for k, v in model.state_dict().items()
            if k.startswith(prefix) and "frequency_bands" not in k}


def bundle_dict(b, tag):
    return {f"{tag}_rgb": b.rgb_map, f"{tag}_depth": b.depth_map, f"{tag}_weights": b.weights,
            f"{tag}_mask_weights": b.mask_weights, f"{tag}_acc": b.acc_map, f"{tag}_disp": b.disp_map}


def object_rays(H, W, n, seed):
    """n ray ids in the central half of the image (through the object), fixed seed."""
    g = torch.Generator().manual_seed(seed)
    r = torch.randint(H // 4, 3 * H // 4, (n,), generator=g)
    c = torch.randint(W // 4, 3 * W // 4, (n,), generator=g)
    return r * W + c


@torch.no_grad()
def main():
    # ---------------------------------------------------------------- a1/a2 ray generation
    pose = torch.from_numpy(pose_spherical(30.0, -30.0, 4.0))
    o_s, d_s = nerf.get_ray_bundle(12, 10, 13.5, pose)
    o_n, d_n = nerf.ndc_rays(12, 10, 13.5, 1.0, o_s[None, None, :].expand(12, 10, 3), d_s)
    poses120 = np.stack([pose_spherical(a, -30.0, 4.0) for a in np.linspace(-270, 90, 120, endpoint=False)])
    save("golden_raygen.npz", pose=pose, H=12, W=10, focal=13.5, origin=o_s, dirs=d_s,
         ndc_o=o_n, ndc_d=d_n, poses120=poses120[[0, 17, 59, 119]], poses120_angles=np.linspace(-270, 90, 120, endpoint=False)[[0, 17, 59, 119]])

    # ---------------------------------------------------------------- lego NeRF (C2)
    m = rh.load_model("NeRFModel", "colab-lego-nerf-high-res")
    save("weights_lego_nerf.npz", **{f"coarse.{k}": v for k, v in net_state(m, "model_coarse.").items()},
         **{f"fine.{k}": v for k, v in net_state(m, "model_fine.").items()},
         sample_pdf_u=m.sample_pdf.u)
    H = W = 800
    focal = 0.5 * 800 / np.tan(0.5 * 0.6911112)
    o, d = nerf.get_ray_bundle(H, W, float(focal), pose)
    ids = object_rays(H, W, 96, 1)
    dirs = d.view(-1, 3)[ids].contiguous()
    bounds = torch.tensor([2.0, 6.0])
    cb, fb = m.forward((o, dirs, bounds))
    # intermediates, re-derived with the reference's own modules (same calls as model_nerf.py:52-75)
    near, far = bounds
    t_c = m.sampler(m.cfg.nerf.validation, dirs.shape[0], near, far)
    p_c = models.intervals_to_ray_points(t_c, dirs, o)
    raw_c = m.model_coarse(p_c, dirs[..., None, :].expand_as(p_c))
    t_f = m.sample_pdf(t_c, cb.weights, m.cfg.nerf.validation.perturb)
    p_f = models.intervals_to_ray_points(t_f, dirs, o)
    raw_f = m.model_fine(p_f, dirs[..., None, :].expand_as(p_f))
    g = torch.Generator().manual_seed(7)
    pts = (torch.rand(1024, 3, generator=g) * 2 - 1) * 1.2
    pdirs = torch.nn.functional.normalize(torch.randn(1024, 3, generator=g), dim=-1)
    save("golden_lego_nerf.npz", focal=float(focal), H=H, W=W, pose=pose, ray_ids=ids, origin=o, dirs=dirs,
         bounds=bounds, t_coarse=t_c, raw_coarse=raw_c, t_fine=t_f, raw_fine=raw_f,
         **bundle_dict(cb, "coarse"), **bundle_dict(fb, "fine"),
         pts=pts, pdirs=pdirs, sample_points_fine=m.sample_points(pts, pdirs),
         sample_points_coarse=m.model_coarse(pts, pdirs))

    # ---------------------------------------------------------------- grid sweep + iso (C3, small res)
    class A:  # the argparse namespace mesh_nerf.py builds (:205-267)
        limit, res, iso_level, batch_size = 1.2, 20, 32.0, 4096
    rad = mesh_nerf.extract_radiance(m, A, "cpu", A.res)
    iso = mesh_nerf.extract_iso_level(rad[..., 3], A)
    save("golden_lego_grid.npz", limit=A.limit, res=A.res, iso_level=A.iso_level, radiance=rad, iso_value=np.float32(iso),
         lin=torch.linspace(-A.limit, A.limit, A.res))

    # ---------------------------------------------------------------- lego BuFF (C5)
    mb = rh.load_model("BuFFModel", "buff-synthetic-lego")
    save("weights_lego_buff.npz", **{f"coarse.{k}": v for k, v in net_state(mb, "model.").items()},
         voxels=mb.tree.voxels)
    idsb = torch.cat([object_rays(H, W, 80, 2), torch.tensor([0, 799, 5 * 800 + 3, 639999])])  # + corner rays that miss
    dirsb = d.view(-1, 3)[idsb].contiguous()
    bb = mb.forward((o[None], dirsb, bounds))
    z, _, mask = mb.tree.batch_ray_voxel_intersect(o[None], dirsb, near, far, samples_count=192)
    t_u = mb.sampler(mb.cfg.nerf.validation, dirsb.shape[0], near, far)
    z[~mask] = t_u[~mask]
    p_b = models.intervals_to_ray_points(z, dirsb, o[None])
    raw_b = mb.model(p_b, dirsb[..., None, :].expand_as(p_b))
    save("golden_lego_buff.npz", focal=float(focal), H=H, W=W, pose=pose, ray_ids=idsb, origin=o, dirs=dirsb, bounds=bounds,
         z=z, ray_mask=mask, raw=raw_b, **bundle_dict(bb, "out"))

    # ---------------------------------------------------------------- fern NeRF, NDC rays (C4)
    mf = rh.load_model("NeRFModel", "nerf-colmap-fern")
    save("weights_fern_nerf.npz", **{f"coarse.{k}": v for k, v in net_state(mf, "model_coarse.").items()},
         **{f"fine.{k}": v for k, v in net_state(mf, "model_fine.").items()},
         sample_pdf_u=mf.sample_pdf.u)
    Hf, Wf, ff = 756, 1008, 815.13
    posef = torch.eye(4)
    posef[0, 3] = 0.1
    of, df = nerf.get_ray_bundle(Hf, Wf, ff, posef)
    on, dn = nerf.ndc_rays(Hf, Wf, ff, 1.0, of[None, None, :].expand(Hf, Wf, 3), df)
    idf = object_rays(Hf, Wf, 64, 3)
    on_s, dn_s = on.reshape(-1, 3)[idf].contiguous(), dn.reshape(-1, 3)[idf].contiguous()
    bf = torch.tensor([0.0, 1.0])
    cbf, fbf = mf.forward((on_s, dn_s, bf))
    t_cf = mf.sampler(mf.cfg.nerf.validation, 64, bf[0], bf[1])
    t_ff = mf.sample_pdf(t_cf, cbf.weights, mf.cfg.nerf.validation.perturb)
    p_ff = models.intervals_to_ray_points(t_ff, dn_s, on_s)
    raw_ff = mf.model_fine(p_ff, dn_s[..., None, :].expand_as(p_ff))
    save("golden_fern_nerf.npz", focal=ff, H=Hf, W=Wf, pose=posef, ray_ids=idf, origins=on_s, dirs=dn_s, bounds=bf,
         t_coarse=t_cf, t_fine=t_ff, raw_fine=raw_ff, **bundle_dict(cbf, "coarse"), **bundle_dict(fbf, "fine"))

    # ---------------------------------------------------------------- stage-level vectors on synthetic inputs
    g = torch.Generator().manual_seed(11)
    t_in = torch.sort(torch.rand(8, 64, generator=g) * 4 + 2, dim=-1).values
    w_in = torch.rand(8, 64, generator=g) ** 4
    w_in[3] = 0.0                       # all-zero weights -> uniform pdf
    w_in[4, 10:] = 0.0                  # spike
    sp = nerf.SamplePDF(128)
    t_out = sp(t_in, w_in, False)
    raw_in = torch.randn(8, 64, 4, generator=g) * torch.tensor([1.0, 1.0, 1.0, 30.0])
    raw_in[..., :3] = torch.sigmoid(raw_in[..., :3])
    dd = torch.randn(8, 3, generator=g) * 1.5
    vr = nerf.VolumeRenderer(0.2, 0.0, False, attenuation_threshold=1e-5).eval()
    vb = vr(raw_in, t_in, dd)
    vrw = nerf.VolumeRenderer(0.2, 0.0, True, attenuation_threshold=1e-5).eval()
    vbw = vrw(raw_in, t_in, dd)
    rs = nerf.RaySampleInterval(64)

    class C:
        lindisp, perturb = True, False
    t_lindisp = rs(C, 5, torch.tensor(2.0), torch.tensor(6.0))
    C.lindisp = False
    t_perray = rs(C, 5, torch.linspace(0.5, 1.5, 5), torch.linspace(3.0, 7.0, 5))
    pe10 = nerf.PositionalEncoding(10, True, True)(pts[:16] * 5.0)
    pe4 = nerf.PositionalEncoding(4, True, True)(pdirs[:16])
    save("golden_stages.npz", pdf_t=t_in, pdf_w=w_in, pdf_out=t_out, vr_raw=raw_in, vr_t=t_in, vr_dirs=dd,
         **bundle_dict(vb, "vr"), **bundle_dict(vbw, "vrw"), t_lindisp=t_lindisp, t_perray=t_perray,
         pe_in_xyz=pts[:16] * 5.0, pe_xyz=pe10, pe_in_dir=pdirs[:16], pe_dir=pe4)


if __name__ == "__main__":
    main()
