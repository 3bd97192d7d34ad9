"""Golden vectors for the BuFF tree maintenance, produced by the UNMODIFIED reference
(src/nerf/tree.py via ref_harness.py) in this container:  python tests/golden/make_golden_tree.py
  * per-sample voxel indices of batch_ray_voxel_intersect on the golden BuFF rays        (tree.py:215-343)
  * memm after one and two ray_batch_integration calls                                    (tree.py:177-206)
  * voxel lists of a small synthetic tree before / after two consolidate() calls          (tree.py:127-175)
  * the random sampling branch (use_random_sampling) under torch.manual_seed(4321), 48 samples (tree.py:280-297)
"""
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402


def main():
    torch.manual_seed(0)
    rh.install()
    from nerf import tree as rtree
    g = dict(np.load(os.path.join(HERE, "golden_lego_buff.npz")))
    mb = rh.load_model("BuFFModel", "buff-synthetic-lego")
    o, dirs = torch.from_numpy(g["origin"]), torch.from_numpy(g["dirs"])
    near, far = torch.tensor(float(g["bounds"][0])), torch.tensor(float(g["bounds"][1]))
    with torch.no_grad():
        z, idx, mask = mb.tree.batch_ray_voxel_intersect(o[None], dirs, near, far, samples_count=192)
        assert torch.equal(z[mask], torch.from_numpy(g["z"])[mask])
        w, mw = torch.from_numpy(g["out_weights"]), torch.from_numpy(g["out_mask_weights"])
        t = mb.tree
        t.memm = torch.zeros(t.voxels.shape[0])
        t.counter = 1
        step = t.config.tree.step_size_integration_offset
        t.ray_batch_integration(step, idx[mask], w[mask], mw[mask])
        memm1 = t.memm.clone()
        t.ray_batch_integration(step + 1, idx[mask], (w * 0.5)[mask], mw[mask])
        memm2 = t.memm.clone()
        t.ray_batch_integration(step - 5, idx[mask], w[mask], mw[mask])          # before the offset: no-op
        assert torch.equal(t.memm, memm2) and t.counter == 3

        # the random branch (tree.py:280-297) under a fixed global seed: pins the oracle's restatement draw for draw
        mb.tree.config.tree.use_random_sampling = True
        torch.manual_seed(4321)
        z_r, idx_r, mask_r = mb.tree.batch_ray_voxel_intersect(o[None], dirs, near, far, samples_count=48)
        mb.tree.config.tree.use_random_sampling = False
        assert torch.equal(mask_r, mask)

        cfg = NS(dataset=NS(near=2.0, far=6.0),
                 tree=NS(subdivision_outer_count=3, subdivision_inner_count=2, max_depth=3, eps=0.3, max_voxel_count=60,
                         use_random_sampling=False, step_size_integration_offset=10, step_size_tree=4))
        ts = rtree.TreeSampling(cfg, "cpu")
        v0 = ts.voxels.clone()
        gen = torch.Generator().manual_seed(3)
        m1 = torch.rand(v0.shape[0], generator=gen)
        ts.memm = m1.clone()
        ts.consolidate()
        v1 = ts.voxels.clone()
        m2 = torch.rand(v1.shape[0], generator=gen)
        ts.memm = m2.clone()
        ts.consolidate()
        v2 = ts.voxels.clone()
        ticks = np.array([int(ts.ticked(s)) for s in range(0, 30)], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "golden_tree.npz"), idx=idx.numpy().astype(np.int32), ray_mask=mask.numpy(),
                        memm1=memm1.numpy(), memm2=memm2.numpy(), v0=v0.numpy(), m1=m1.numpy(), v1=v1.numpy(), m2=m2.numpy(),
                        v2=v2.numpy(), ticks=ticks, z_random=z_r.numpy(), idx_random=idx_r.numpy().astype(np.int32))
    print("idx", idx.shape, "memm nonzero", int((memm1 != 0).sum()), "voxels", v0.shape[0], v1.shape[0], v2.shape[0])


if __name__ == "__main__":
    main()
