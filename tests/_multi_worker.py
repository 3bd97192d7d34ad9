"""Worker of tests/test_gpu_multi.py — launched with torchrun, one process per GPU (NCCL).  Every rank renders its
row shard / computes its grid slab, the exchange assembles the result, and every rank compares it with the single-GPU
result it computes itself: the assembled image must be BIT-IDENTICAL (no arithmetic crosses a shard boundary) and the
assembled mesh must equal the single-GPU mesh array for array.  Prints `MULTI_OK <world>` from rank 0 on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_npz  # noqa: E402
from test_gpu_parity import BUFF_CFG, LEGO_CFG  # noqa: E402


def main():
    import nerfmeshes_b200 as nm
    from nerfmeshes_b200 import parallel as par
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = load_npz("golden_lego_nerf.npz")
    pose = g["pose"]

    # ---- row-sharded lego image (H not divisible by the world size on purpose) + the one all_gather
    lego = nm.NeRFModel.from_npz(LEGO_CFG, load_npz("weights_lego_nerf.npz")).eval().cuda()
    eng = lego._engine()
    assert eng.device.index == local, "engine must follow the parameters' device"
    for (H, W, f) in ((45, 64, 70.0), (48, 40, 55.0)):
        want = ("rgb", "depth", "acc", "disp")
        full = eng.render_image(pose, H, W, f, 2.0, 6.0, want=list(want))
        got = par.render_image_sharded(lego, pose, H, W, f, 2.0, 6.0, want=want)
        for k in want:
            assert torch.equal(got[k], full[k]), f"lego {k} {H}x{W}: gathered image differs from the single-GPU image"

    # ---- NDC rays (fern weights), per-ray origins generated on every rank
    fern = nm.NeRFModel.from_npz(LEGO_CFG, load_npz("weights_fern_nerf.npz")).eval().cuda()
    H, W, f = 42, 56, 60.0
    full = fern._engine().render_image(torch.eye(4), H, W, f, 0.0, 1.0, ndc=True, want=["rgb", "disp"])
    got = par.render_image_sharded(fern, torch.eye(4), H, W, f, 0.0, 1.0, ndc=True, want=("rgb", "disp"))
    assert torch.equal(got["rgb"], full["rgb"]) and torch.equal(got["disp"], full["disp"]), "fern NDC shards differ"

    # ---- BuFF (AABB-clipped sampling; voxel list replicated)
    buff = nm.BuFFModel.from_npz(BUFF_CFG, load_npz("weights_lego_buff.npz")).eval().cuda()
    beng = buff._engine()
    buff._sync_tree(beng)
    H, W, f = 40, 40, 55.0
    full = beng.render_image(pose, H, W, f, 2.0, 6.0, buff=True, want=["rgb", "acc"])
    got = par.render_image_sharded(buff, pose, H, W, f, 2.0, 6.0, buff=True, want=("rgb", "acc"))
    assert torch.equal(got["rgb"], full["rgb"]) and torch.equal(got["acc"], full["acc"]), "BuFF shards differ"

    # ---- slab-sharded mesh: sigma sweep + iso statistics + marching cubes + gather == the single-GPU arrays
    class Args:
        res, limit, iso_level = 40, 1.2, 32.0
    v1, f1, n1, iso1 = par.extract_geometry_sharded(lego, Args, group=par.SINGLE, to_host=False)
    vN, fN, nN, isoN = par.extract_geometry_sharded(lego, Args, to_host=False)
    assert iso1 == isoN, (iso1, isoN)
    assert v1.shape[0] > 100 and f1.shape[0] > 100
    assert torch.equal(vN, v1), "gathered vertex array differs from the single-GPU one"
    assert torch.equal(fN, f1), "gathered face array differs from the single-GPU one"
    assert torch.equal(nN, n1), "gathered normals differ (halo planes)"

    dist.barrier()
    if rank == 0:
        print(f"MULTI_OK {world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
