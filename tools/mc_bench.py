"""Stage timing of the marching-cubes path on one GPU (host wall clock around each library call + CUDA events), on a
synthetic 512^3 field (a bumpy sphere: ~1.5 M vertices, like the lego grid) or a .npy volume.

    python tools/mc_bench.py [--res 512] [--volume density.npy] [--iso 0.0] [--reps 5]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def lego(a):
    """wall-clock and CUDA-event stage times of the full mesh path on the lego checkpoint."""
    import nerfmeshes_b200 as nm
    from nerfmeshes_b200 import parallel as par
    from bench import load_npz, model_cfg
    model = nm.NeRFModel.from_npz(model_cfg(2.0, 6.0), load_npz("weights_lego_nerf.npz")).eval().cuda()
    eng = model._engine()
    wall = {}

    def wrap(name):
        fn = getattr(eng, name)

        def timed(*args, **kw):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(*args, **kw)
            torch.cuda.synchronize()
            wall[name] = wall.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return out
        setattr(eng, name, timed)
    for name in ("grid_sigma", "volume_stats", "mc_count", "mc_emit"):
        wrap(name)

    class Args:
        res, limit, iso_level = a.res, 1.2, 32.0
    for rep in range(a.reps):
        wall.clear()
        tm = {}
        t0 = time.perf_counter()
        v, f, n, iso = par.extract_geometry_sharded(model, Args, to_host=False, timings=tm)
        torch.cuda.synchronize()
        print(f"rep {rep}: total wall {1e3 * (time.perf_counter() - t0):.2f} ms | events " +
              ", ".join(f"{k} {x:.3f}" for k, x in tm.items()) + " | wall per call " + ", ".join(f"{k} {x:.3f}" for k, x in wall.items()) +
              f" | {v.shape[0]} vertices {f.shape[0]} triangles")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--volume")
    ap.add_argument("--iso", type=float, default=0.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--lego", action="store_true", help="time parallel.extract_geometry_sharded (lego sigma sweep + MC) stage by stage")
    a = ap.parse_args()
    if a.lego:
        return lego(a)
    from nerfmeshes_b200.nerf_api import _engine
    eng = _engine()
    if a.volume:
        vol = torch.from_numpy(np.load(a.volume).astype(np.float32)).cuda()
    else:
        g = torch.linspace(-1.2, 1.2, a.res, device="cuda")
        X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
        vol = (0.8 + 0.05 * torch.sin(9 * X) * torch.sin(7 * Y) * torch.sin(11 * Z) - torch.sqrt(X * X + Y * Y + Z * Z)).contiguous()
        del X, Y, Z
    n0 = vol.shape[0]
    for rep in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nv, nt = eng.mc_count(vol, a.iso, 0, n0, 0, n0)
        t1 = time.perf_counter()
        verts = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
        normals = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
        faces = torch.empty((nt, 3), dtype=torch.int32, device="cuda")
        t2 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.mc_emit(vol, a.iso, 0, n0, 0, n0, nv, nt, 0, out=(verts, normals, faces))
        e1.record()
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        print(f"rep {rep}: {nv} vertices {nt} triangles | count (4 kernels + readback) {1e3 * (t1 - t0):.3f} ms wall | alloc {1e3 * (t2 - t1):.3f} ms | "
              f"emit launch {1e3 * (t3 - t2):.3f} ms wall, {e0.elapsed_time(e1):.3f} ms device | sync {1e3 * (t4 - t3):.3f} ms | "
              f"total {1e3 * (t4 - t0):.3f} ms -> {4 * vol.numel() / (t4 - t0) / 1e9:.0f} GB/s of the one-read volume")
        del verts, normals, faces


if __name__ == "__main__":
    main()
