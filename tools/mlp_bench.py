"""Micro-benchmark of the fused-MLP kernel alone (points resident in HBM): TFLOP/s per precision / debug switch.

    python tools/mlp_bench.py [--tiles-per-sm 64]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(prec, M, reps, sigma_only=False):
    import torch
    import nerfmeshes_b200 as nm
    from oracle import nerf_oracle as O
    cfg = O.NetCfg()
    sd = O.init_weights(cfg, 1)
    eng = nm.Engine(cfg.__dict__, None, nm.RenderSettings(num_coarse=8, num_fine=0, precision=prec))
    eng.load_weights(0, sd)
    g = torch.Generator(device="cuda").manual_seed(0)
    pts = (torch.rand(M, 3, device="cuda", generator=g) * 2 - 1) * 1.2
    dirs = torch.randn(M, 3, device="cuda", generator=g)
    try:
        for _ in range(2):
            eng.point_mlp(0, pts, dirs, sigma_only=sigma_only)
        torch.cuda.synchronize()
    except Exception as e:
        print("MLPBENCH " + json.dumps(dict(error=str(e)[:80], flags=eng.kernel_flags(), dbg=os.environ.get("NM_TC_DEBUG", "0"),
                                            stages=os.environ.get("NM_TC_STAGES", "max"))), flush=True)
        os._exit(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.point_mlp(0, pts, dirs, sigma_only=sigma_only)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = cfg.flops_per_point(sigma_only) * M
    return dict(prec=prec, M=M, ms=ms, tflops=flop / ms / 1e9, pts_per_s=M / ms * 1e3, dbg=os.environ.get("NM_TC_DEBUG", "0"),
                stages=os.environ.get("NM_TC_STAGES", "max"), sigma_only=sigma_only)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles-per-sm", type=int, default=64)
    ap.add_argument("--one", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    M = 148 * 128 * a.tiles_per_sm
    if a.one is not None:
        prec, so = a.one.split(",")
        print("MLPBENCH " + json.dumps(run(int(prec), M, 3, so == "1")), flush=True)
        sys.exit(0)
    combos = [("0", "0,0"), ("0", "1,0"), ("0", "0,1"), ("1", "0,0"), ("2", "0,0"), ("4", "0,0"), ("3", "0,0"), ("6", "0,0"), ("7", "0,0"),
              ("1", "1,0"), ("2", "1,0")]
    if a.quick:
        combos = [("0", "0,0"), ("0", "1,0"), ("0", "0,1"), ("1", "0,0"), ("3", "0,0")]
    for dbg, one in combos:
        env = dict(os.environ, NM_TC_DEBUG=dbg)
        r = subprocess.run([sys.executable, __file__, "--tiles-per-sm", str(a.tiles_per_sm), "--one", one], env=env,
                           capture_output=True, text=True, timeout=300)
        out = [l for l in r.stdout.splitlines() if l.startswith("MLPBENCH")]
        print(out[-1] if out else f"MLPBENCH fail dbg={dbg} {one}: {r.stderr[-300:]}", flush=True)
    for ns in (("2", "3") if a.quick else ("2", "3", "4")):
        env = dict(os.environ, NM_TC_DEBUG="0", NM_TC_STAGES=ns)
        r = subprocess.run([sys.executable, __file__, "--tiles-per-sm", str(a.tiles_per_sm), "--one", "0,0"], env=env,
                           capture_output=True, text=True, timeout=300)
        out = [l for l in r.stdout.splitlines() if l.startswith("MLPBENCH")]
        print(out[-1] if out else f"MLPBENCH fail stages={ns}: {r.stderr[-300:]}", flush=True)
