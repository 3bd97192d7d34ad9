"""Energy per issued tensor FLOP of the fused MLP kernel against cuBLAS bf16 GEMM, both run long enough to sit at the
board's power cap ("a power trace showing the cap binds at equal energy per FLOP").

Samples nvidia-smi (power.draw, clocks.sm, sw_power_cap) every 100 ms while (a) torch.matmul bf16 8192^3 and (b) full
800x800 lego images (64+128 samples, exact mode: 3 tensor-core products per algorithmic product) run back to back for
`--seconds` each, and prints one JSON object: rate (TFLOP/s, issued), median power, J per issued TFLOP, median SM clock.
A tool for evidence — torch.matmul is the yardstick here, not part of the product path.

    python tools/power_trace.py [--seconds 8] > profiles/r02_power_trace.json"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Sampler:
    def __init__(self):
        self.rows = []
        self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap,power.limit",
                                      "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def window(self, t0, t1):
        r = [x for t, x in self.rows if t0 <= t <= t1]
        pw = [float(x[0]) for x in r]
        return {"samples": len(r), "power_w_median": float(np.median(pw)) if pw else None,
                "sm_mhz_median": float(np.median([float(x[1]) for x in r])) if r else None,
                "power_capped_frac": (sum(x[2].lower() == "active" for x in r) / len(r)) if r else None,
                "power_limit_w": float(r[0][3]) if r else None,
                "trace_w": [round(p) for p in pw]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=8.0)
    a = ap.parse_args()
    import nerfmeshes_b200 as nm
    from bench import load_npz, model_cfg, FLOP_PER_POINT
    smp = Sampler()
    out = {}
    # (a) cuBLAS bf16
    n = 8192
    A = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    B = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(20):
        A @ B
    torch.cuda.synchronize()
    time.sleep(1.0)
    t0 = time.perf_counter()
    it = 0
    while time.perf_counter() - t0 < a.seconds:
        for _ in range(50):
            A @ B
        torch.cuda.synchronize()
        it += 50
    t1 = time.perf_counter()
    w = smp.window(t0 + 1.0, t1)       # skip the ramp
    rate = it * 2 * n ** 3 / (t1 - t0) / 1e12
    out["cublas_bf16_8192"] = {**w, "tflops_issued": rate, "joule_per_issued_tflop": w["power_w_median"] / rate if w["power_w_median"] else None}
    time.sleep(2.0)
    # (b) the fused MLP kernel inside full images
    model = nm.NeRFModel.from_npz(model_cfg(2.0, 6.0), load_npz("weights_lego_nerf.npz")).eval().cuda()
    eng = model._engine()
    pose = torch.tensor([[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 4.0]])      # any pose: the work per ray is pose-independent
    focal = 1111.1
    for _ in range(2):
        eng.render_image(pose, 800, 800, focal, 2.0, 6.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    it = 0
    while time.perf_counter() - t0 < a.seconds:
        eng.render_image(pose, 800, 800, focal, 2.0, 6.0)
        torch.cuda.synchronize()
        it += 1
    t1 = time.perf_counter()
    w = smp.window(t0 + 1.0, t1)
    alg = it * 640000 * (64 + 192) * FLOP_PER_POINT / (t1 - t0) / 1e12
    out["mlp_tc_kernel_lego_800"] = {**w, "tflops_algorithmic": alg, "tflops_issued": 3 * alg,
                                     "joule_per_issued_tflop": w["power_w_median"] / (3 * alg) if w["power_w_median"] else None}
    smp.proc.terminate()
    a_, b_ = out["cublas_bf16_8192"], out["mlp_tc_kernel_lego_800"]
    if a_["joule_per_issued_tflop"] and b_["joule_per_issued_tflop"]:
        out["energy_per_issued_flop_vs_cublas"] = b_["joule_per_issued_tflop"] / a_["joule_per_issued_tflop"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
