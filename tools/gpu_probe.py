"""Staged bring-up probe for the fused-MLP kernels on a real B200 (run under gpurun).

Each stage runs in its own subprocess with a timeout, from the simplest network that exercises one mechanism to the
full 8x256 net, and prints max-abs errors against the CPU oracle.  A hang or trap in one stage cannot take the
others down.  Usage:  python tools/gpu_probe.py [stage ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = {
    # name: (net kwargs, precision, M)
    "simt_full":      (dict(num_layers=8, hidden_size=256, num_encoding_fn_xyz=10, num_encoding_fn_dir=4), 2, 1000),
    "tc_L1_noview":   (dict(num_layers=1, hidden_size=128, num_encoding_fn_xyz=4, use_viewdirs=False), 0, 128),
    "tc_L2_noview":   (dict(num_layers=2, hidden_size=128, num_encoding_fn_xyz=4, use_viewdirs=False), 0, 128),
    "tc_L2_256":      (dict(num_layers=2, hidden_size=256, num_encoding_fn_xyz=10, use_viewdirs=False), 0, 300),
    "tc_tiny":        (dict(num_layers=4, hidden_size=128, num_encoding_fn_xyz=6, num_encoding_fn_dir=4), 0, 1000),
    "tc_full":        (dict(num_layers=8, hidden_size=256, num_encoding_fn_xyz=10, num_encoding_fn_dir=4), 0, 5000),
    "tc_full_fast":   (dict(num_layers=8, hidden_size=256, num_encoding_fn_xyz=10, num_encoding_fn_dir=4), 1, 5000),
    "tc_full_big":    (dict(num_layers=8, hidden_size=256, num_encoding_fn_xyz=10, num_encoding_fn_dir=4), 0, 148 * 128 * 9 + 77),
}


def run_stage(name):
    import torch
    from oracle import nerf_oracle as O
    from nerfmeshes_b200.engine import Engine, RenderSettings
    kw, prec, M = STAGES[name]
    cfg = O.NetCfg(**{**dict(num_layers=4, hidden_size=128, skip_step=4, num_encoding_fn_xyz=6, num_encoding_fn_dir=4), **kw})
    sd = O.init_weights(cfg, seed=5)
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(M, 3, generator=g) * 2 - 1) * 1.5
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    with torch.no_grad():
        ref = O.flexible_nerf_forward(sd, cfg, pts, dirs)
    eng = Engine(cfg.__dict__, None, RenderSettings(num_coarse=8, num_fine=0, precision=prec))
    eng.load_weights(0, sd)
    out = eng.point_mlp(0, pts.cuda(), dirs.cuda())
    torch.cuda.synchronize()
    out = out.cpu()
    err = (out - ref).abs()
    res = dict(stage=name, M=M, rgb_max=float(err[:, :3].max()), sigma_max=float(err[:, 3].max()),
               sigma_ref_absmax=float(ref[:, 3].abs().max()), nan=int(torch.isnan(out).sum()),
               first_bad_row=int((err.max(1).values > 1e-2).nonzero()[0]) if bool((err.max(1).values > 1e-2).any()) else -1,
               frac_bad_rows=float((err.max(1).values > 1e-2).float().mean()))
    if res["frac_bad_rows"] > 0:
        bad = (err.max(1).values > 1e-2).nonzero().flatten()[:8].tolist()
        res["bad_rows"] = bad
        res["sample"] = [[round(float(x), 5) for x in out[b]] + [round(float(x), 5) for x in ref[b]] for b in bad[:3]]
    sg = eng.point_mlp(0, pts.cuda(), dirs.cuda(), sigma_only=True).cpu()
    res["sigma_only_max"] = float((sg - ref[:, 3]).abs().max())
    print("PROBE " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        run_stage(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--stage", n], capture_output=True, text=True, timeout=180)
            lines = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            print(lines[-1] if lines else f"PROBE {{\"stage\": \"{n}\", \"rc\": {r.returncode}, \"stderr\": {json.dumps(r.stderr[-600:])}}}", flush=True)
        except subprocess.TimeoutExpired:
            print(f'PROBE {{"stage": "{n}", "timeout": true}}', flush=True)
