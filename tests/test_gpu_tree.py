"""BuFF tree maintenance on the device: nm_ray_voxel_indices / nm_tree_integrate against the oracle and
the reference-generated goldens, and the training-mode hook of BuFFModel.forward."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, load_npz
from oracle import nerf_oracle as O
from test_gpu_parity import BUFF_CFG

pytestmark = pytest.mark.gpu
G = dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_tree.npz")))


@pytest.fixture(scope="module")
def buff():
    import nerfmeshes_b200 as nm
    return nm.BuFFModel.from_npz(BUFF_CFG, load_npz("weights_lego_buff.npz")).cuda().eval()


def test_voxel_indices_match_oracle_bit_exact(buff):
    g = load_npz("golden_lego_buff.npz")
    vox = load_npz("weights_lego_buff.npz")["voxels"].float()
    near, far = float(g["bounds"][0]), float(g["bounds"][1])
    eng = buff._engine()
    buff._sync_tree(eng)
    idx, z = eng.ray_voxel_indices(g["origin"][None].cuda(), g["dirs"].cuda(), near, far, want_z=True)
    z_ref, idx_ref, mask = O.batch_ray_voxel_intersect(vox, g["origin"][None], g["dirs"], near, far, 192, return_indices=True)
    idx, z = idx.cpu(), z.cpu()
    assert torch.equal(z[mask], z_ref[mask]) and torch.equal(z, g["z"])            # placement incl. the uniform fallback rows
    assert bool((idx[~mask] == -1).all()) and torch.equal(idx[mask], idx_ref[mask].int())


def test_tree_integration_matches_reference(buff):
    g = load_npz("golden_lego_buff.npz")
    eng = buff._engine()
    mask = torch.from_numpy(G["ray_mask"])
    idx = torch.from_numpy(G["idx"]).clone()
    idx[~mask] = -1                                                    # whole batch, misses flagged (the reference passes x[mask])
    w, mw = g["out_weights"].cuda(), g["out_mask_weights"].cuda()
    memm = torch.zeros(buff.tree.voxels.shape[0], device="cuda")
    eng.tree_integrate(idx.cuda(), w, mw, memm, 1)
    assert float((memm.cpu() - torch.from_numpy(G["memm1"])).abs().max()) < 1e-6
    eng.tree_integrate(idx.cuda()[mask.cuda()], (w * 0.5)[mask.cuda()], mw[mask.cuda()], memm, 2)     # ...[mask] rows work too
    assert float((memm.cpu() - torch.from_numpy(G["memm2"])).abs().max()) < 1e-6


def test_training_forward_feeds_the_tree_and_consolidate_rebuilds_the_voxels():
    import nerfmeshes_b200 as nm
    g = load_npz("golden_lego_buff.npz")
    cfg = {**BUFF_CFG, "nerf.train.radiance_field_noise_std": 0.0, "tree.step_size_integration_offset": 5, "tree.step_size_tree": 3,
           "tree.eps": 1e-4, "tree.max_depth": 4, "tree.subdivision_inner_count": 2, "tree.max_voxel_count": 1536}
    model = nm.BuFFModel.from_npz(cfg, load_npz("weights_lego_buff.npz")).cuda().train()
    rays = (g["origin"][None].cuda(), g["dirs"].cuda(), g["bounds"])
    model.global_step = 2
    with torch.no_grad():
        model.forward(rays)
    assert model.tree.counter == 1                                      # before the offset: nothing accumulated
    model.global_step = 5
    with torch.no_grad():
        out = model.forward(rays)
        model.forward(rays)
    assert model.tree.counter == 3 and model.tree.memm.is_cuda
    hit = torch.from_numpy(G["ray_mask"])
    vox = load_npz("weights_lego_buff.npz")["voxels"].float()
    _, idx, _ = O.batch_ray_voxel_intersect(vox, g["origin"][None], g["dirs"], float(g["bounds"][0]), float(g["bounds"][1]), 192,
                                            return_indices=True)
    ref, c = torch.zeros(vox.shape[0]), 1
    for _ in range(2):
        ref, c = O.ray_batch_integration(ref, c, idx[hit], out.weights.cpu()[hit], out.mask_weights.cpu()[hit])
    assert float((model.tree.memm.cpu() - ref).abs().max()) < 1e-5 and int((ref > 0).sum()) > 20
    # graft a node graph onto the flat checkpoint voxels, then prune + subdivide and render with the new list
    from nerfmeshes_b200.tree import Node
    model.tree.root.children = [Node(model.tree.config, (b[0].clone(), b[1].clone()), 3) for b in vox]
    n_before = vox.shape[0]
    model.tree.consolidate()
    assert model.tree.voxels.shape[0] != n_before and model.tree.counter == 1 and float(model.tree.memm.abs().sum()) == 0
    with torch.no_grad():
        out2 = model.forward(rays)
    assert bool(torch.isfinite(out2.rgb_map).all())


def test_buff_training_step_updates_weights_tree_and_schedule():
    """nerfmeshes_b200.training_step on a BuFFModel (model_buff.py:75-110): loss + gradients, weight integration into the
    tree on every step past the offset, consolidate() when the schedule ticks."""
    import nerfmeshes_b200 as nm
    from nerfmeshes_b200.tree import Node
    g = load_npz("golden_lego_buff.npz")
    z = load_npz("weights_lego_buff.npz")
    cfg = {**BUFF_CFG, "nerf.train.radiance_field_noise_std": 0.0, "tree.step_size_integration_offset": 2, "tree.step_size_tree": 2,
           "tree.eps": 1e-4, "tree.max_depth": 4, "tree.subdivision_inner_count": 2, "tree.max_voxel_count": 1536}
    model = nm.BuFFModel.from_npz(cfg, z).cuda().train()
    model.tree.root.children = [Node(model.tree.config, (b[0].clone(), b[1].clone()), 3) for b in z["voxels"].float()]
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    rays = (g["origin"][None].cuda(), g["dirs"].cuda(), g["bounds"])
    target = torch.rand(g["dirs"].shape[0], 3, generator=torch.Generator().manual_seed(0)).cuda()
    n0 = model.tree.voxels.shape[0]
    sizes, losses = [], []
    for step in range(5):
        out = nm.training_step(model, rays, target, global_step=step)
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in model.parameters())
        opt.step()
        losses.append(out["loss"])
        sizes.append((model.tree.voxels.shape[0], model.tree.counter))
    # steps 0,1: before the offset; step 2,3 integrate; step 4 = offset + 2 ticks -> integrate then consolidate (counter back to 1)
    assert sizes[0] == (n0, 1) and sizes[1] == (n0, 1) and sizes[2] == (n0, 2) and sizes[3] == (n0, 3)
    assert sizes[4][1] == 1 and sizes[4][0] != n0
    assert abs(out["log"]["train/psnr"] + 10 * np.log10(out["loss"])) < 1e-6 and np.isfinite(losses).all()


def test_random_voxel_sampling_branch(buff):
    """cfg.tree.use_random_sampling (src/nerf/tree.py:280-297): multinomial voxel draws with replacement, a uniform depth
    inside the drawn voxel's [entry, exit], then the common sort.  torch's generator stream cannot be matched, so the parity
    is distributional and structural: every sample lies inside the interval of the voxel it is attributed to, that voxel is
    one the ray hits (the oracle's hit mask), samples ascend, rays without a hit keep the uniform fallback / index -1, every
    hit voxel of a ray is drawn about S/H times, the render with the flag uses exactly these samples, and a new seed draws
    new ones."""
    g = load_npz("golden_lego_buff.npz")
    vox = load_npz("weights_lego_buff.npz")["voxels"].float()
    near, far = float(g["bounds"][0]), float(g["bounds"][1])
    eng = buff._engine()
    buff._sync_tree(eng)
    o, d = g["origin"][None], g["dirs"]
    _, _, mask = O.batch_ray_voxel_intersect(vox, o, d, near, far, 192, return_indices=True)
    eng.voxel_random = True
    try:
        idx, z = eng.ray_voxel_indices(o.cuda(), d.cuda(), near, far, want_z=True, seed=11)
        idx2, z2 = eng.ray_voxel_indices(o.cuda(), d.cuda(), near, far, want_z=True, seed=12)
        out = eng.render_rays(o.cuda(), d.cuda(), near, far, buff=True, seed=11, want=["rgb", "t_vals"])
    finally:
        eng.voxel_random = False
    idx, z = idx.cpu(), z.cpu()
    assert torch.equal(out["t_vals"].cpu(), z)                           # the render's own samples
    assert bool((idx[~mask] == -1).all()) and torch.equal(z[~mask], g["z"][~mask])
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    assert not torch.equal(z2.cpu()[mask], z[mask])
    # slab intersection per (ray, voxel) in float64 — interval membership with a few ulps of slack
    inv = 1.0 / d.double()
    a0 = (vox[None, :, 0].double() - o.double()[:, None]) * inv[:, None]     # (R,V,3)
    a1 = (vox[None, :, 1].double() - o.double()[:, None]) * inv[:, None]
    lo_c, hi_c = torch.minimum(a0, a1), torch.maximum(a0, a1)
    t_in, t_out = lo_c.max(-1).values, hi_c.min(-1).values                 # (R,V)
    hit = (t_in <= t_out) & (t_in >= near) & (t_out <= far)
    rows = torch.nonzero(mask).flatten()
    for r in rows.tolist():
        v = idx[r].long()
        assert bool((v >= 0).all()) and bool(hit[r, v].all())
        assert bool((z[r].double() >= t_in[r, v] - 1e-5).all()) and bool((z[r].double() <= t_out[r, v] + 1e-5).all())
        H = int(hit[r].sum())
        counts = torch.bincount(v, minlength=vox.shape[0])[hit[r]]
        assert int(counts.sum()) == 192
        if H <= 24:                                                       # expected 192/H >= 8 per voxel: none may starve
            assert int(counts.min()) >= 1 and int(counts.max()) <= 192 // H * 4 + 8
    assert torch.isfinite(out["rgb"]).all()
