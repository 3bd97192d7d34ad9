"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard arithmetic and the gather / statistics exchange.
The per-shard computation itself needs a GPU; here shards are filled with synthetic data."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, ROOT)
from nerfmeshes_b200 import parallel as P  # noqa: E402


def test_shard_arithmetic():
    for H in (800, 756, 7, 1):
        for world in (1, 2, 3, 8):
            spans = [P.row_shard(H, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    for n0 in (512, 20, 9):
        for world in (1, 2, 4, 8):
            slabs = [P.slab_shard(n0, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == n0
            assert all(a[1] - 1 == b[0] for a, b in zip(slabs, slabs[1:]))      # exactly one shared plane
            assert sum(e - s - 1 for s, e in slabs) == n0 - 1                      # every cell layer exactly once


def test_slab_layout_owns_every_plane_once():
    for n0 in (512, 40, 9):
        for world in (1, 2, 4, 8):
            if n0 - 1 < world:
                continue
            lay = [P.slab_layout(n0, r, world) for r in range(world)]
            assert lay[0][0] == 0 and lay[-1][1] == n0
            assert all(a[1] == b[0] for a, b in zip(lay, lay[1:]))              # owned point planes tile [0, n0)
            for own0, own1, buf0, buf1 in lay:
                assert buf0 == max(own0 - 1, 0) and buf1 == min(own1 + 2, n0)   # one halo plane below, two above


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W = 9, 4
        full = torch.arange(H * W * 3, dtype=torch.float32).reshape(H * W, 3)
        r0, r1 = P.row_shard(H, rank, world)
        got = P.gather_rows({"rgb": full[r0 * W:r1 * W], "acc": full[r0 * W:r1 * W, 0]})
        ok = torch.equal(got["rgb"], full) and torch.equal(got["acc"], full[:, 0])
        # mesh soup: rank r contributes r+2 vertices and r+1 faces indexing its own vertices
        v = torch.full((rank + 2, 3), float(rank))
        f = torch.arange(rank + 1, dtype=torch.int32)[:, None].repeat(1, 3)
        V, F, N = P.gather_mesh(v, f, v.clone())
        ok &= V.shape[0] == sum(r + 2 for r in range(world)) and torch.equal(N, V)
        off = sum(r + 2 for r in range(rank))
        mine = F[sum(r + 1 for r in range(rank)):sum(r + 1 for r in range(rank + 1))]
        ok &= torch.equal(mine, f + off) and bool((V[mine.long()] == float(rank)).all())
        # sharded statistics equal the single-array ones
        vol = torch.from_numpy(np.random.default_rng(0).standard_normal((11, 5, 6)).astype(np.float32) * 700)
        x0, x1 = P.slab_shard(11, rank, world)
        sl = vol[x0:x1]
        own = sl if rank == world - 1 else sl[:-1]
        mn, mx, sd = P.global_stats(float(own.min()), float(own.max()), float(own.double().sum()),
                                    lambda m: float(((own.double() - m) ** 2).sum()), own.numel(), torch.device("cpu"))
        ok &= mn == float(vol.min()) and mx == float(vol.max()) and abs(sd - float(vol.double().std(unbiased=False))) < 1e-9
        # the row exchange: kernels fill `views`, ONE all_gather_into_tensor assembles every map (H not divisible by world)
        for H2 in (9, 8):
            ex = P.RowExchange(torch.device("cpu"), H2, W, ("rgb", "acc"))
            fullrgb = torch.arange(H2 * W * 3, dtype=torch.float32).reshape(H2 * W, 3) + 0.5
            ex.views["rgb"].copy_(fullrgb[ex.r0 * W:ex.r1 * W])
            ex.views["acc"].copy_(fullrgb[ex.r0 * W:ex.r1 * W, 1])
            got = ex.gather()
            ok &= torch.equal(got["rgb"], fullrgb) and torch.equal(got["acc"], fullrgb[:, 1])
        # halo planes of an x-slab arrive from the neighbours
        n0 = 11
        own0, own1, buf0, buf1 = P.slab_layout(n0, rank, world)
        buf = torch.full((buf1 - buf0, 5, 6), -1.0)
        buf[own0 - buf0:own1 - buf0] = vol[own0:own1]
        P.exchange_halo_planes(buf, n0, rank, world)
        ok &= torch.equal(buf, vol[buf0:buf1])
        # data-parallel gradient exchange: mean over ranks, one flat all_reduce, parameters without a gradient skipped
        ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1))]
        ps[0].grad = torch.full((3, 2), float(rank + 1))
        ps[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
        n = P.allreduce_gradients(ps)
        mean = sum(r + 1 for r in range(world)) / world
        ok &= n == 11 and torch.equal(ps[0].grad, torch.full((3, 2), mean)) and ps[2].grad is None
        ok &= torch.allclose(ps[1].grad, torch.arange(5, dtype=torch.float32) * mean)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_and_stats_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
