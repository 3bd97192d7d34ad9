"""Multi-GPU sharding of the hot path: one process per GPU, independent units, ONE exchange step.

Render: contiguous image-row shards, every rank generates its own rays from the 12-float pose; results are bit-identical
to the single-GPU image (no cross-shard arithmetic), the exchange is one all_gather of the finished rows.
Mesh: x-slabs of the density grid; every grid point (vertex, cell) is owned by exactly one rank; halo planes by send/recv;
per-slab marching cubes in global index coordinates with globally consistent vertex ids; the exchange is an all_gather
of the vertex / triangle counts and ONE all_gather of the per-slab indexed meshes (padded to the largest slab), plus
scalar all_reduces for the iso-level statistics (extract_iso_level, src/mesh_nerf.py:56-65).  The gathered arrays equal
the single-GPU arrays bit for bit.
Training: data parallel over rays — every rank runs forward + backward on its own ray batch (no collective inside), then
ONE all_reduce of the flattened gradients of both networks (595 k - 1.19 M floats, 4.8 MB) before the optimiser step.
The collectives go through torch.distributed (NCCL on GPUs, gloo in the CPU tests); there is no data-path collective
inside a shard's computation.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def row_shard(H: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [r0, r1) of rank `rank`: H split as evenly as possible, earlier ranks take the remainder."""
    base, rem = divmod(H, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def slab_shard(n0: int, rank: int, world: int) -> Tuple[int, int]:
    """Planes [x0, x1) of the grid's slowest axis owned by `rank` such that the n0-1 cell layers are split evenly and
    neighbouring slabs share one plane (cells between planes x1-1 and x1 belong to the next rank's first plane)."""
    c0, c1 = row_shard(n0 - 1, rank, world)          # cell layers [c0, c1)
    return c0, c1 + 1                                # planes c0 .. c1 inclusive


def _all_gather_padded(t: torch.Tensor, group=None):
    """all_gather of tensors whose first dimension differs per rank -> list of per-rank tensors."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:c] for o, c in zip(out, counts)]


def gather_rows(local: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Concatenate per-rank row shards (rank order == row order) of every output map."""
    return {k: torch.cat(_all_gather_padded(v.contiguous(), group), 0) for k, v in local.items()}


def gather_mesh(verts: torch.Tensor, faces: torch.Tensor, normals: Optional[torch.Tensor] = None, group=None):
    """Concatenate per-slab indexed meshes; face indices are shifted by the exclusive scan of the vertex counts.
    Vertices on shared planes appear once per adjacent slab with identical bits (use torch.unique for the set)."""
    vs = _all_gather_padded(verts.contiguous(), group)
    fs = _all_gather_padded(faces.contiguous(), group)
    ns = _all_gather_padded(normals.contiguous(), group) if normals is not None else None
    off, shifted = 0, []
    for v, f in zip(vs, fs):
        shifted.append(f + off)
        off += v.shape[0]
    return torch.cat(vs, 0), torch.cat(shifted, 0), (torch.cat(ns, 0) if ns is not None else None)


def global_stats(local_min: float, local_max: float, local_sum: float, local_sumsq_centered_fn, count: int, device, group=None):
    """min / max / population std of a sharded volume: two rounds (mean first, then centred squares) so the result
    matches the two-pass single-GPU computation."""
    t = torch.tensor([local_min, -local_max], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    s = torch.tensor([local_sum, float(count)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    mean = float(s[0] / s[1])
    q = torch.tensor([local_sumsq_centered_fn(mean)], dtype=torch.float64, device=device)
    dist.all_reduce(q, op=dist.ReduceOp.SUM, group=group)
    return float(t[0]), float(-t[1]), math.sqrt(float(q[0]) / float(s[1]))


def allreduce_gradients(params, group=None, average=True):
    """Data-parallel training step, exchange part: one all_reduce over the flattened .grad of `params` (the parameters of
    both FlexibleNeRFModels); the mean over ranks is what a single process would get from the concatenated ray batch
    when every rank's loss is a mean over equally many rays (src/models/model_nerf.py:118-126)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return int(flat.numel())


class RowExchange:
    """The exchange step of a row-sharded image: one flat buffer per rank, laid out map after map
    ([rgb | depth | acc | disp], every segment padded to the largest shard), which the render kernels fill through
    `views`, and ONE `all_gather_into_tensor` (NCCL over NVLink) that leaves every full map on every rank."""
    MAPS = ("rgb", "depth", "depth_raw", "acc", "disp", "coarse_rgb", "coarse_acc", "coarse_disp")

    def __init__(self, device, H, W, want=("rgb", "depth", "acc", "disp"), group=None):
        self.group, self.H, self.W, self.want = group, H, W, tuple(want)
        assert all(k in self.MAPS for k in self.want), "only per-ray maps can be exchanged"
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.r0, self.r1 = row_shard(H, self.rank, self.world)
        self.widths = {k: (3 if k.endswith("rgb") else 1) for k in self.want}
        self.rmax = (H + self.world - 1) // self.world * W            # rays of the largest shard
        C = sum(self.widths.values())
        self.local = torch.empty(C * self.rmax, dtype=torch.float32, device=device)
        self.full = torch.empty(self.world * C * self.rmax, dtype=torch.float32, device=device)
        R, off = (self.r1 - self.r0) * W, 0
        self.views = {}
        for k in self.want:
            w = self.widths[k]
            self.views[k] = self.local[off:off + w * R].view((R, 3) if w == 3 else (R,))
            off += w * self.rmax

    def gather(self):
        dist.all_gather_into_tensor(self.full, self.local, group=self.group)
        full = self.full.view(self.world, -1)
        out, off = {}, 0
        n = self.H * self.W
        for k in self.want:
            w = self.widths[k]
            if self.H % self.world == 0:
                seg = full[:, off:off + w * self.rmax].reshape((n, 3) if w == 3 else (n,))
            else:
                spans = [row_shard(self.H, r, self.world) for r in range(self.world)]
                parts = [full[r, off:off + w * (b - a) * self.W] for r, (a, b) in enumerate(spans)]
                seg = torch.cat(parts).view((n, 3) if w == 3 else (n,))
            out[k] = seg
            off += w * self.rmax
        return out


_EXCHANGES = {}


def row_exchange(device, H, W, want, group=None) -> RowExchange:
    key = (str(device), H, W, tuple(want), id(group))
    if key not in _EXCHANGES:
        _EXCHANGES.clear()
        _EXCHANGES[key] = RowExchange(device, H, W, want, group)
    return _EXCHANGES[key]


def render_image_sharded(model, pose, H, W, focal, near, far, *, ndc=False, buff=False, want=("rgb", "depth", "acc", "disp"),
                         group=None, seed=0):
    """eval_nerf.py's image loop on N GPUs: this rank renders image rows [r0, r1) — rays generated on the device
    from the pose, kernels writing straight into this rank's segment of the exchange buffer — then ONE all_gather assembles
    every output map on every rank.  No arithmetic crosses a shard boundary, so the assembled maps are bit-identical to
    the single-GPU image."""
    eng = model._engine()
    if buff:
        model._sync_tree(eng)
    ex = row_exchange(eng.device, H, W, want, group)
    eng.render_image(pose, H, W, focal, near, far, ndc=ndc, rows=(ex.r0, ex.r1), buff=buff, seed=seed, want=list(ex.want),
                     out=ex.views)
    return ex.gather()


_MESH_BUFFERS = {}


def _scratch(key, numel, dtype, device):
    """Grow-only scratch tensors of the mesh path (slab buffer, exchange segments, single-GPU outputs): steady-state calls
    make no allocator traffic.  Tensors returned by extract_geometry_sharded(to_host=False) are views of these buffers
    and stay valid until the next call."""
    t = _MESH_BUFFERS.get((key, str(device)))
    if t is None or t.numel() < numel or t.dtype != dtype:
        t = torch.empty(int(numel * 1.25) + 16, dtype=dtype, device=device)
        _MESH_BUFFERS[(key, str(device))] = t
    return t[:numel]


SINGLE = "single"        # pass as `group` to run the sharded code paths as ONE shard (tests compare it with the N-rank result)


def _rank_world(group=None):
    if group is SINGLE or group == SINGLE:
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class _StageTimer:
    """CUDA-event stage timing on the current stream (filled into a caller-supplied dict as `<stage>_ms`)."""

    def __init__(self, sink):
        self.sink, self.marks = sink, []

    def mark(self, name=None):
        if self.sink is None:
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.marks.append((name, e))

    def finish(self):
        if self.sink is None:
            return
        torch.cuda.current_stream().synchronize()
        for (_, a), (name, b) in zip(self.marks, self.marks[1:]):
            self.sink[name + "_ms"] = self.sink.get(name + "_ms", 0.0) + a.elapsed_time(b)


def slab_layout(n0: int, rank: int, world: int):
    """x-slab of rank `rank` of a grid with n0 planes: returns (own0, own1, buf0, buf1).
    Points (and the cells above them) of planes [own0, own1) are OWNED: the n0-1 cell layers are split evenly and the last
    rank also owns the top plane n0-1.  Marching cubes additionally needs, where they exist, plane own1 (upper corners of
    the last owned cell layer), own1+1 and own0-1 (central-difference normals): the buffer is planes [buf0, buf1)."""
    c0, c1 = row_shard(n0 - 1, rank, world)
    own1 = n0 if rank == world - 1 else c1
    return c0, own1, max(c0 - 1, 0), min(own1 + 2, n0)


def exchange_halo_planes(buf: torch.Tensor, n0: int, rank: int, world: int, group=None):
    """Fill the halo planes of `buf` (planes [buf0,buf1) of slab_layout) from the neighbouring ranks: this rank's first two
    owned planes go down to rank-1, its last owned plane goes up to rank+1 — three 4*n1*n2-byte planes per interior rank
    over NVLink (NCCL send/recv), instead of recomputing 3 planes of MLP per rank.  Needs >= 2 owned planes on every rank."""
    own0, own1, buf0, buf1 = slab_layout(n0, rank, world)
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, buf[own0 - buf0:own0 - buf0 + 2], rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, buf[0:own0 - buf0], rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, buf[own1 - 1 - buf0:own1 - buf0], rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, buf[own1 - buf0:buf1 - buf0], rank + 1, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


def extract_geometry_sharded(model, args, group=None, to_host=True, timings=None, halo="exchange"):
    """mesh_nerf.extract_geometry (src/mesh_nerf.py:68-92) on N GPUs, x-slabs of the grid:
      1. every rank sweeps sigma over ITS planes (fused MLP, grid front-end) into its slab buffer;
      2. halo planes: 3 planes per interior rank by send/recv from the neighbours (`halo="exchange"`), or recomputed;
      3. extract_iso_level: min / max / std over the whole grid — scalar all_reduces (every plane is owned exactly once);
      4. marching cubes, count step; all_gather of the (n_vertices, n_triangles) pairs -> every rank's index offset;
      5. marching cubes, emit step, straight into this rank's segment of the exchange buffer; ONE all_gather of
         [vertices | normals | faces] (padded to the largest slab) leaves the whole mesh on every rank.
    A vertex belongs to the rank that owns its grid point, and the last owned cell layer addresses the next rank's
    vertices by the ids that rank assigns (nm_mc_count / nm_mc_emit), so the concatenation IS the single-GPU mesh — same
    arrays, bit for bit; there are no duplicates to remove.  Returns (vertices, triangles, normals, iso) like the single-GPU
    function (vertices rescaled to (-limit, limit) when to_host).  Works without a process group (one slab)."""
    import numpy as np
    rank, world = _rank_world(group)
    eng = model._engine()
    res, dev = args.res, eng.device
    tm = _StageTimer(timings)
    tiles = [torch.linspace(-args.limit, args.limit, res) for _ in range(3)]
    own0, own1, buf0, buf1 = slab_layout(res, rank, world)
    tm.mark()
    buf = _scratch("slab", (buf1 - buf0) * res * res, torch.float32, dev).view(buf1 - buf0, res, res)
    can_exchange = world > 1 and halo == "exchange" and all(
        (lambda a: a[1] - a[0] >= 2)(slab_layout(res, r, world)) for r in range(world))
    if world == 1 or can_exchange:
        eng.grid_sigma(tiles, own0, own1, out=buf[own0 - buf0:own1 - buf0])
        tm.mark("sweep")
        if world > 1:
            exchange_halo_planes(buf, res, rank, world, group)
    else:
        eng.grid_sigma(tiles, buf0, buf1, out=buf)            # halo planes recomputed (bit-identical to the owner's)
        tm.mark("sweep")
    own = buf[own0 - buf0:own1 - buf0]
    tm.mark("halo")
    if world > 1:
        # two-pass statistics like the single-GPU nm_volume_stats, shards combined on the device: two tiny all_gathers, and one
        # host read at the end (every plane is owned exactly once, so sums are exact partitions)
        acc = torch.empty(5, dtype=torch.float64, device=dev)
        acc[4] = float(own.numel())
        eng.volume_stats_pass(own, 1, acc)
        allacc = torch.empty((world, 5), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allacc, acc, group=group)
        mean = (allacc[:, 2].sum() / allacc[:, 4].sum()).reshape(1)
        eng.volume_stats_pass(own, 2, acc, mean)
        allsq = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allsq, acc[3:4].contiguous(), group=group)
        st3 = torch.stack([allacc[:, 0].min(), allacc[:, 1].max(), (allsq.sum() / allacc[:, 4].sum()).sqrt()]).cpu()
        smin, smax, sstd = (float(np.float32(float(x))) for x in st3)
    else:
        smin, smax, sstd = eng.volume_stats(own)
    iso = float(min(max(args.iso_level, np.float32(smin) + np.float32(sstd)), np.float32(smax) - np.float32(sstd)))
    tm.mark("stats")
    shard = (iso, buf0, res, own0 - buf0, own1 - buf0)
    nv, nt = eng.mc_count(buf, *shard)
    if world == 1:
        outs = (_scratch("v", 3 * nv, torch.float32, dev).view(nv, 3), _scratch("n", 3 * nv, torch.float32, dev).view(nv, 3),
                _scratch("f", 3 * nt, torch.int32, dev).view(nt, 3))
        v, f, n = eng.mc_emit(buf, *shard, nv, nt, 0, out=outs)
        tm.mark("mc")
        tm.mark("gather")
    else:
        counts = torch.tensor([nv, nt], dtype=torch.int64, device=dev)
        allc = torch.empty((world, 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, counts, group=group)
        allc = allc.cpu()
        nvs, nts = [int(x) for x in allc[:, 0]], [int(x) for x in allc[:, 1]]
        v_base = sum(nvs[:rank])
        if sum(nvs) >= 2 ** 31:
            raise OverflowError("mesh too large for int32 indices")
        vmax, tmax = max(max(nvs), 1), max(max(nts), 1)
        seg = 3 * (2 * vmax + tmax)                               # floats per rank: vertices | normals | faces (int32 bits)
        local = _scratch("local", seg, torch.float32, dev)
        full = _scratch("full", world * seg, torch.float32, dev)
        views = (local[:3 * vmax].view(vmax, 3), local[3 * vmax:6 * vmax].view(vmax, 3), local[6 * vmax:].view(torch.int32).view(tmax, 3))
        eng.mc_emit(buf, *shard, nv, nt, v_base, out=views)
        tm.mark("mc")
        dist.all_gather_into_tensor(full, local, group=group)
        full = full.view(world, seg)
        v = torch.cat([full[r, :3 * nvs[r]] for r in range(world)]).view(-1, 3)
        n = torch.cat([full[r, 3 * vmax:3 * vmax + 3 * nvs[r]] for r in range(world)]).view(-1, 3)
        f = torch.cat([full[r, 6 * vmax:6 * vmax + 3 * nts[r]] for r in range(world)]).view(torch.int32).view(-1, 3)
        tm.mark("gather")
    tm.finish()
    if to_host:
        v = args.limit * (v.cpu() / (res / 2.0) - 1.0)
        return v, f.cpu(), n.cpu(), iso
    return v, f, n, iso
