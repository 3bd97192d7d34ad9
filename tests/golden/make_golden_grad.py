"""Golden vector for the training backward, produced by the UNMODIFIED reference in this container:
    python tests/golden/make_golden_grad.py
The reference's NeRFModel (shipped lego checkpoint, validation-mode sampling: no jitter / noise, so the run is deterministic)
renders 48 golden rays, loss = mse(coarse.rgb_map, target) + mse(fine.rgb_map, target) as in training_step
(src/models/model_nerf.py:118-126), loss.backward() through the reference's own modules.  Stored per parameter tensor: L2 norm,
sum, and 24 entries at fixed pseudo-random positions — enough to pin torch autograd on oracle/nerf_oracle.py (the oracle the CUDA
backward is tested against) to the reference's gradients without committing 4.8 MB of them."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402


def probe_index(numel, k=24, seed=5):
    return torch.from_numpy(np.random.RandomState(seed + numel % 9973).randint(0, numel, size=k)).long()


def main():
    torch.manual_seed(0)
    rh.install()
    g = dict(np.load(os.path.join(HERE, "golden_lego_nerf.npz")))
    m = rh.load_model("NeRFModel", "colab-lego-nerf-high-res")
    m.eval()                                              # cfg.nerf.validation: perturb False, noise 0 (model_nerf.py:48)
    R = 48
    o, dirs = torch.from_numpy(g["origin"]), torch.from_numpy(g["dirs"])[:R].contiguous()
    bounds = torch.from_numpy(g["bounds"])
    target = torch.rand(R, 3, generator=torch.Generator().manual_seed(11))
    for p in m.parameters():
        p.grad = None
    coarse, fine = m.forward((o, dirs, bounds))
    lc = torch.nn.functional.mse_loss(coarse.rgb_map, target)
    lf = torch.nn.functional.mse_loss(fine.rgb_map, target)
    (lc + lf).backward()
    out = {"R": R, "target": target.numpy(), "loss_coarse": lc.item(), "loss_fine": lf.item()}
    n = 0
    for name, p in m.named_parameters():
        if not (name.startswith("model_coarse.") or name.startswith("model_fine.")) or p.grad is None:
            continue
        key = name.replace("model_coarse.", "coarse.").replace("model_fine.", "fine.")
        gflat = p.grad.detach().flatten().double()
        idx = probe_index(gflat.numel())
        out[f"{key}|norm"] = float(gflat.norm())
        out[f"{key}|sum"] = float(gflat.sum())
        out[f"{key}|probe"] = gflat[idx].numpy()
        n += 1
    np.savez_compressed(os.path.join(HERE, "golden_lego_grad.npz"), **out)
    print("tensors", n, "loss", lc.item(), lf.item())

    # BuFF: single network on the AABB-clipped samples (model_buff.py:34-69), loss = mse(rgb_map, target) (:96-104)
    gb = dict(np.load(os.path.join(HERE, "golden_lego_buff.npz")))
    mb = rh.load_model("BuFFModel", "buff-synthetic-lego")
    mb.eval()
    ob, db = torch.from_numpy(gb["origin"]), torch.from_numpy(gb["dirs"])
    tb = torch.rand(db.shape[0], 3, generator=torch.Generator().manual_seed(12))
    for p in mb.parameters():
        p.grad = None
    bundle = mb.forward((ob[None], db, torch.from_numpy(gb["bounds"])))
    lb = torch.nn.functional.mse_loss(bundle.rgb_map, tb)
    lb.backward()
    outb = {"target": tb.numpy(), "loss": lb.item()}
    nb = 0
    for name, p in mb.named_parameters():
        if p.grad is None:
            continue
        gflat = p.grad.detach().flatten().double()
        idx = probe_index(gflat.numel())
        outb[f"{name}|norm"] = float(gflat.norm())
        outb[f"{name}|sum"] = float(gflat.sum())
        outb[f"{name}|probe"] = gflat[idx].numpy()
        nb += 1
    np.savez_compressed(os.path.join(HERE, "golden_buff_grad.npz"), **outb)
    print("buff tensors", nb, "loss", lb.item(), [k for k in outb if k.endswith("|norm")][:4])


if __name__ == "__main__":
    main()
