#!/usr/bin/env python
"""bench.py — rays/sec of the NeRF render hot path on N B200s (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload lego|buff|mesh]

One step = one pass of the hot path over one batch of synthetic input: an 800x800 image (640,000 rays, 64 coarse + 128
fine samples, two 8x256 MLPs) rendered from one of the 120 `SynthesizableDataset.synthesis` poses
(src/data/datasets.py:105-130) with the weights of the reference's shipped lego checkpoint (re-packed as
tests/golden/weights_lego_nerf.npz; no dataset or network needed).  Prints ONE JSON line (rank 0).

  value     rays/s, rays generated on the device from the pose (inputs resident), CUDA-event timed, max over ranks
  e2e       rays/s through the host-buffer C-ABI call (model.query with CPU tensors): ray directions H2D from pinned
            memory and rgb/disp D2H inside the timed region, every step
  roofline  the fused-MLP kernel against the measured bf16 tensor peak (algorithmic FLOPs: 1,186,816 per point)
  cpu_baseline / --impl reference: the oracle port (torch-CPU restatement of the reference, same ATen kernels the
            reference itself runs) on the host cores, bounded sample
N > 1: one process per GPU (torchrun), every rank renders its own images (independent units, no data-path
collective): weak scaling; value = all ranks' rays / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_POINT = 1186816
H = W = 800
FOCAL = 0.5 * 800 / np.tan(0.5 * 0.6911112)
NEAR, FAR = 2.0, 6.0
NC, NF = 64, 128


def load_npz(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    return {k: torch.from_numpy(z[k]) for k in z.files if z[k].dtype.kind == "f"}


def lego_cfg(buff=False):
    from oracle.nerf_oracle import NetCfg
    net = NetCfg().__dict__
    cfg = {"experiment.model": "NeRFModel", "dataset.near": NEAR, "dataset.far": FAR, "dataset.white_background": False,
           "models.coarse_type": "FlexibleNeRFModel", "models.fine_type": "FlexibleNeRFModel", "models.use_fine": not buff,
           **{f"models.coarse.{k}": v for k, v in net.items()}, **{f"models.fine.{k}": v for k, v in net.items()}}
    for mode in ("train", "validation"):
        cfg.update({f"nerf.{mode}.num_coarse": 192 if buff else NC, f"nerf.{mode}.num_fine": NF, f"nerf.{mode}.perturb": False,
                    f"nerf.{mode}.lindisp": False, f"nerf.{mode}.radiance_field_noise_std": 0.0})
    if buff:
        cfg["tree.subdivision_outer_count"] = 2
    return cfg


def poses120():
    from oracle.nerf_oracle import pose_spherical
    return [pose_spherical(float(a), -30.0, 4.0) for a in np.linspace(-270, 90, 120, endpoint=False)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1407.1), d.get("hbm_gbs", 6564.5), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(steps, warmup, chunk=2048):
    """The reference's CPU implementation of the path, restated (oracle/nerf_oracle.py calls the same ATen ops in the
    same order): NeRFModel.query on `chunk`-ray batches (the shipped validation chunksize) of the lego workload."""
    from oracle import nerf_oracle as O
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    z = load_npz("weights_lego_nerf.npz")
    coarse = {k[7:]: v for k, v in z.items() if k.startswith("coarse.")}
    fine = {k[5:]: v for k, v in z.items() if k.startswith("fine.")}
    net, rc = O.NetCfg(), O.RenderCfg()
    pose = poses120()[40]
    o, d = O.get_ray_bundle(H, W, float(FOCAL), pose)
    d = d.reshape(-1, 3)
    # "all the host threads it can use": intra-op scaling of 256-wide GEMMs saturates early and oversubscribed boxes get
    # slower with more threads, so probe a few thread counts on a short chunk and keep the fastest
    best = (None, 1e30)
    with torch.no_grad():
        for nt in sorted({avail, max(avail // 2, 1), 32, 16, 8} & set(range(1, avail + 1))):
            torch.set_num_threads(nt)
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                O.nerf_forward(coarse, fine, net, net, rc, o, d[320200:320200 + 512], torch.tensor(NEAR), torch.tensor(FAR), u=z["sample_pdf_u"])
                ts.append(time.perf_counter() - t0)
            if min(ts) < best[1]:
                best = (nt, min(ts))
    cores = best[0]
    torch.set_num_threads(cores)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            r0 = 320000 + 200 + (i % 4) * chunk                     # centre-of-image chunks (through the object)
            t0 = time.perf_counter()
            O.nerf_forward(coarse, fine, net, net, rc, o, d[r0:r0 + chunk], torch.tensor(NEAR), torch.tensor(FAR), u=z["sample_pdf_u"])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    tot = sum(times)
    return (chunk * steps / tot, cores, f"{steps} x {chunk}-ray chunks of the lego 800x800 64+128 workload, torch {torch.__version__} "
            f"CPU, {cores} threads (fastest of a probe over thread counts; {avail} logical CPUs visible)", tot / steps * 1e3)


def cpu_train_run(cores, rays=256, steps=2):
    """The reference's training step on the CPU (oracle forward + torch autograd backward, model_nerf.py:88-151)."""
    from oracle import nerf_oracle as O
    z = load_npz("weights_lego_nerf.npz")
    leaf = lambda d: {k: (v.clone().requires_grad_(True) if k.endswith((".weight", ".bias")) else v) for k, v in d.items()}
    coarse = leaf({k[7:]: torch.as_tensor(v) for k, v in z.items() if k.startswith("coarse.")})
    fine = leaf({k[5:]: torch.as_tensor(v) for k, v in z.items() if k.startswith("fine.")})
    net, rc = O.NetCfg(), O.RenderCfg()
    o, d = O.get_ray_bundle(H, W, float(FOCAL), poses120()[40])
    d = d.reshape(-1, 3)[320000:320000 + rays]
    tgt = torch.rand(rays, 3, generator=torch.Generator().manual_seed(0))
    torch.set_num_threads(cores)
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        bc, bf, _, _ = O.nerf_forward(coarse, fine, net, net, rc, o, d, torch.tensor(NEAR), torch.tensor(FAR), u=torch.as_tensor(z["sample_pdf_u"]))
        (torch.nn.functional.mse_loss(bc.rgb_map, tgt) + torch.nn.functional.mse_loss(bf.rgb_map, tgt)).backward()
        if i:
            times.append(time.perf_counter() - t0)
    return rays * steps / sum(times), f"{steps} x {rays}-ray forward+backward steps of the same workload, torch autograd on {cores} CPU threads"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="exact", choices=["exact", "fast", "fp32"])
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "lego synthetic 800x800, 64 coarse + 128 fine samples, 8x256 MLP x2 (configs[1])",
              "rays_per_step_per_gpu": H * W, "poses": "SynthesizableDataset.synthesis (120, r=4, phi=-30)",
              "weights": "pretrained/colab-lego-nerf-high-res re-packed (tests/golden/weights_lego_nerf.npz)",
              "parallelism": f"image-parallel x{world} (independent poses per rank, no collective)",
              "l2": "per-step working set ~3 GB of samples >> 126 MB L2; no explicit flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return
        v, cores, sample, ms = cpu_reference_run(a.steps, min(a.warmup, 1))
        print(json.dumps({"impl": "reference", "metric": "rays/sec", "value": v, "unit": "rays/s", "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import nerfmeshes_b200 as nm
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = nm.NeRFModel.from_npz(lego_cfg(), load_npz("weights_lego_nerf.npz")).eval()
    model.precision = {"exact": nm.PREC_EXACT, "fast": nm.PREC_FAST, "fp32": nm.PREC_FP32}[a.precision]
    model.cuda(local)
    eng = model._engine()
    poses = poses120()
    my_pose = lambda i: poses[(i * world + rank) % len(poses)]
    want = ["rgb", "depth", "acc", "disp"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident arm
    for i in range(a.warmup):
        eng.render_image(my_pose(i), H, W, FOCAL, NEAR, FAR, want=want)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    eng.set_timing(True)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        out = eng.render_image(my_pose(a.warmup + i), H, W, FOCAL, NEAR, FAR, want=want)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    mlp_ms, mlp_pts, mlp_n = eng.mlp_time_ms()
    eng.set_timing(False)
    clk = clocks.stop() if rank == 0 else None
    finite = bool(torch.isfinite(out["rgb"]).all())

    # ---------------------------------------------------------------- end-to-end arm (host buffers through the C ABI)
    from oracle.nerf_oracle import get_ray_bundle
    o_h, d_h = get_ray_bundle(H, W, float(FOCAL), poses[0])
    d_h = d_h.reshape(-1, 3).contiguous().pin_memory()
    o_h = o_h.contiguous().pin_memory()
    host_want = ["rgb", "disp"]
    for i in range(2):
        eng.render_rays(o_h, d_h, NEAR, FAR, want=host_want)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        q = eng.render_rays(o_h, d_h, NEAR, FAR, want=host_want)     # nm_query_host: H2D dirs, render, D2H rgb+disp, sync
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---------------------------------------------------------------- secondary: 512^3 sigma sweep + marching cubes (configs[2])
    from nerfmeshes_b200 import parallel as par
    RES, LIMIT, ISO = 512, 1.2, 32.0
    tiles = [torch.linspace(-LIMIT, LIMIT, RES) for _ in range(3)]
    x0, x1 = par.slab_shard(RES, rank, world)
    warm = eng.grid_sigma(tiles, x0, min(x0 + 8, x1))               # warm-up (kernel + workspace allocations)
    eng.marching_cubes(torch.zeros((x1 - x0, RES, RES), device="cuda"), ISO)
    del warm
    barrier()
    g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    g0.record()
    sigma = eng.grid_sigma(tiles, x0, x1)                           # sigma-only fast path (rgb is discarded by the reference, mesh_nerf.py:73)
    gsig = torch.cuda.Event(enable_timing=True)
    gsig.record()
    mv, mf, mn = eng.marching_cubes(sigma, ISO, x_off=float(x0))    # first call sizes torch's output allocations
    del mv, mf, mn
    g1.record()
    mv, mf, mn = eng.marching_cubes(sigma, ISO, x_off=float(x0))
    g2.record()
    barrier()
    grid_ms, mc_ms = g0.elapsed_time(gsig), g1.elapsed_time(g2)
    n_mesh = torch.tensor([mv.shape[0], mf.shape[0]], dtype=torch.float64, device="cuda")
    del sigma, mv, mf, mn

    # ---------------------------------------------------------------- secondary: one training step
    # fused forward + mse(coarse)+mse(fine) + backward of both networks on TRAIN_RAYS centre-of-image rays per rank
    TRAIN_RAYS, TRAIN_STEPS = 4096, 3
    d_tr = d_h[320000:320000 + TRAIN_RAYS].cuda()
    o_tr = o_h.cuda()
    tgt = torch.rand(TRAIN_RAYS, 3, generator=torch.Generator().manual_seed(0)).cuda()
    eng.zero_grad()
    eng.loss_backward(o_tr, d_tr, NEAR, FAR, tgt, training=True, seed=0)          # warm-up (workspace allocation, weight packs)
    barrier()
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l_tr0 = eng.launch_count()
    t0e.record()
    for i in range(TRAIN_STEPS):
        eng.zero_grad()
        loss = eng.loss_backward(o_tr, d_tr, NEAR, FAR, tgt, training=True, seed=1 + i)
    t1e.record()
    barrier()
    train_ms = t0e.elapsed_time(t1e) / TRAIN_STEPS
    train_launches = (eng.launch_count() - l_tr0) // TRAIN_STEPS
    train_loss = [float(x) for x in loss]

    t = torch.tensor([dev_ms, e2e_s * 1e3, grid_ms, mc_ms, train_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_mesh, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, grid_ms, mc_ms, train_ms = (float(x) for x in t)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_mlp_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("mean_bytes_per_launch")       # ncu dram read+write per MLP launch, same workload
    rays = H * W * a.steps * world
    peak_tf, _, peak_src = measured_peaks()
    achieved_tf = (mlp_pts * FLOP_PER_POINT / (mlp_ms * 1e-3)) / 1e12 if mlp_ms > 0 else None
    result = {
        "metric": "rays/sec", "value": rays / (dev_ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"exact": "f16x3 split operands, f32 accumulate", "fast": "f16, f32 accumulate", "fp32": "f32"}[a.precision],
        "data": "synthetic", "config": config, "clocks": clk, "finite": finite,
        "e2e": {"value": rays / (e2e_ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": int(d_h.numel() * 4 + 12),
                "d2h_bytes_per_step": int(H * W * 4 * 4), "api": "nm_query_host (model.query with CPU tensors)"},
        "gpu_launches": int(launches),
        "grid": {"metric": "grid-voxels/sec", "value": RES ** 3 / ((grid_ms + mc_ms) * 1e-3), "unit": "voxels/s", "res": RES,
                 "workload": "lego 512^3 sigma sweep (sigma-only trunk, 982,528 FLOP/voxel) + marching cubes iso=32, x-slabs across ranks",
                 "sigma_sweep_ms": grid_ms, "marching_cubes_ms": mc_ms, "n_vertices": int(n_mesh[0]), "n_triangles": int(n_mesh[1]),
                 "sweep_tflops": RES ** 3 * 982528 / (grid_ms * 1e-3) / 1e12},
        "train": {"metric": "train-rays/sec", "value": TRAIN_RAYS * world / (train_ms * 1e-3), "unit": "rays/s",
                  "rays_per_step_per_gpu": TRAIN_RAYS, "ms_per_step": train_ms, "launches_per_step": int(train_launches),
                  "workload": "nm_loss_backward: fused forward + mse(coarse)+mse(fine) + backward of both 8x256 networks "
                              "(64+192 samples per ray), gradients accumulated on device; no optimiser step",
                  "loss": train_loss,
                  "algorithmic_tflops": TRAIN_RAYS * (64 + 192) * FLOP_PER_POINT * 4 / (train_ms * 1e-3) / 1e12,
                  "frac_of_tensor_peak": TRAIN_RAYS * (64 + 192) * FLOP_PER_POINT * 4 / (train_ms * 1e-3) / 1e12 / peak_tf,
                  "note": "4x forward FLOPs per step: forward, recompute, data gradient, weight gradient; exact mode issues 3 MMAs "
                          "per product, so tensor-pipe work is 3x the algorithmic figure"},
        "roofline": {"bound": "tensor", "kernel": "mlp_tc_kernel", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                     "traffic_note": "DRAM bytes per launch (ncu); algorithmic 20 B/point = 1.64 GB per launch on average: no re-reads",
                     "peak_source": peak_src,
                     "algorithmic_flop_per_point": FLOP_PER_POINT, "points_per_step": mlp_pts // max(a.steps, 1),
                     "launches_per_step": mlp_n // max(a.steps, 1), "kernel_ms_per_step": mlp_ms / max(a.steps, 1),
                     "note": "exact mode issues 3 MMAs per product: tensor-pipe work is 3x the algorithmic FLOPs"},
    }
    if not a.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_run(a.cpu_steps, 1)
        result["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}
        tv, tsample = cpu_train_run(cores)
        result["train"]["cpu_baseline"] = {"value": tv, "unit": "rays/s", "cores": cores, "kind": "port", "sample": tsample}
    print(json.dumps(result))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
