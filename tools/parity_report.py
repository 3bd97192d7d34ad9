"""Parity report on a B200: error of the CUDA path against the committed reference outputs, next to the reference's
own fp32 noise floor (fp32 oracle vs an fp64 evaluation of the same network on the same inputs).

    python tools/parity_report.py [--out gpurun_out/parity.json]
"""
import argparse
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_npz, net_weights  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402
from test_gpu_parity import BUFF_CFG, LEGO_CFG  # noqa: E402

NET = O.NetCfg()


def stats(err):
    e = torch.as_tensor(err).abs().flatten().double()
    q = lambda p: float(torch.quantile(e, p))
    return {"max": float(e.max()), "p99.9": q(0.999), "p99": q(0.99), "median": q(0.5)}


def dbl(sd):
    return {k: v.double() for k, v in sd.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    a = ap.parse_args()
    import nerfmeshes_b200 as nm
    rep = {}
    for name, wfile, gfile in (("lego", "weights_lego_nerf.npz", "golden_lego_nerf.npz"),
                               ("fern", "weights_fern_nerf.npz", "golden_fern_nerf.npz")):
        z, g = load_npz(wfile), load_npz(gfile)
        fine = net_weights(z, "fine")
        model = nm.NeRFModel.from_npz(LEGO_CFG, z).eval()
        eng = model._engine()
        okey = "origin" if name == "lego" else "origins"
        org, dirs = g[okey], g["dirs"]
        near, far = float(g["bounds"][0]), float(g["bounds"][1])
        r = {}
        # ---- per-point: reference raw_fine at the reference's own sample points
        t_f = g["t_fine"]
        p_f = O.intervals_to_ray_points(t_f, dirs, org)
        d_f = dirs[:, None, :].expand_as(p_f)
        with torch.no_grad():
            raw64 = O.flexible_nerf_forward(dbl(fine), NET, p_f.double(), d_f.double()).float()
            raw32 = O.flexible_nerf_forward(fine, NET, p_f, d_f)
        for prec, pname in ((nm.PREC_EXACT, "exact"), (nm.PREC_FP32, "fp32"), (nm.PREC_FAST, "fast")):
            model.precision = prec
            out = model.sample_points(p_f.reshape(-1, 3).cuda(), d_f.reshape(-1, 3).cuda()).cpu().reshape(raw64.shape)
            r[f"point_rgb_{pname}_vs_fp64"] = stats(out[..., :3] - raw64[..., :3])
            r[f"point_sigma_{pname}_vs_fp64"] = stats(out[..., 3] - raw64[..., 3])
            r[f"point_rgb_{pname}_vs_ref"] = stats(out[..., :3] - g["raw_fine"][..., :3])
        r["point_rgb_ref32_vs_fp64"] = stats(g["raw_fine"][..., :3] - raw64[..., :3])
        r["point_sigma_ref32_vs_fp64"] = stats(g["raw_fine"][..., 3] - raw64[..., 3])
        r["point_rgb_oracle32here_vs_fp64"] = stats(raw32[..., :3] - raw64[..., :3])
        r["sigma_range"] = [float(raw64[..., 3].min()), float(raw64[..., 3].max())]
        # ---- teacher-forced composite
        for prec, pname in ((nm.PREC_EXACT, "exact"), (nm.PREC_FP32, "fp32"), (nm.PREC_FAST, "fast")):
            model.precision = prec
            eng = model._engine()
            o = eng.render_rays(org.cuda(), dirs.cuda(), near, far, teacher_t=t_f.cuda(), want=["rgb", "acc", "disp", "depth_raw"])
            r[f"teacher_rgb_{pname}_vs_ref"] = stats(o["rgb"].cpu() - g["fine_rgb"])
            r[f"teacher_acc_{pname}_vs_ref"] = stats(o["acc"].cpu() - g["fine_acc"])
            r[f"teacher_disp_{pname}_vs_ref"] = stats(o["disp"].cpu() - g["fine_disp"])
        b64 = O.volume_render(raw64, t_f, dirs)
        r["teacher_rgb_ref32_vs_fp64net"] = stats(g["fine_rgb"] - b64.rgb_map)
        # ---- end to end
        for prec, pname in ((nm.PREC_EXACT, "exact"), (nm.PREC_FP32, "fp32")):
            model.precision = prec
            c, f = model.forward((org.cuda(), dirs.cuda(), g["bounds"]))
            r[f"e2e_coarse_rgb_{pname}_vs_ref"] = stats(c.rgb_map.cpu() - g["coarse_rgb"])
            r[f"e2e_coarse_w_{pname}_vs_ref"] = stats(c.weights.cpu() - g["coarse_weights"])
            r[f"e2e_fine_rgb_{pname}_vs_ref"] = stats(f.rgb_map.cpu() - g["fine_rgb"])
            r[f"e2e_fine_acc_{pname}_vs_ref"] = stats(f.acc_map.cpu() - g["fine_acc"])
            r[f"e2e_fine_disp_{pname}_vs_ref"] = stats(f.disp_map.cpu() - g["fine_disp"])
            tv = model._engine().render_rays(org.cuda(), dirs.cuda(), near, far, want=["t_vals"])["t_vals"].cpu()
            r[f"e2e_t_fine_{pname}_vs_ref"] = stats(tv - g["t_fine"])
            mask_ref = g["fine_depth"] == 0
            r[f"e2e_depth_mask_disagree_{pname}"] = float(((f.depth_map.cpu() == 0) != mask_ref).float().mean())
        rep[name] = r
        model.precision = nm.PREC_EXACT
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)
    for k, v in rep.items():
        print(k)
        for kk, vv in v.items():
            print("   ", kk, vv)


if __name__ == "__main__":
    main()
