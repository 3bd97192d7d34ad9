"""BuFF voxel tree — host mirror of src/nerf/tree.py:4-212,343-358 (Node, TreeSampling: construction, ticked, consolidate,
ray_batch_integration, flatten, (de)serialize).  The tree itself is a small python object graph that changes every few
thousand steps; what runs per training step — scattering the sample weights of a ray batch into the voxels
(ray_batch_integration) — goes to the library (nm_ray_voxel_indices + nm_tree_integrate).  Pickled `Node` graphs of shipped
checkpoints load into these classes (models._CkptUnpickler), so a resumed tree can be consolidated further."""
from __future__ import annotations

import torch

from . import _lib as L


def _get(node, path, default=None):
    for p in path.split("."):
        if isinstance(node, dict):
            if p not in node:
                return default
            node = node[p]
        elif hasattr(node, p):
            node = getattr(node, p)
        else:
            return default
    return node


class Node:
    """src/nerf/tree.py:4-36.  bounds = (min (3,), max (3,)) fp32; depth 0 is the root (outer subdivision count)."""

    def __init__(self, config=None, bounds=None, depth=0):
        self.config, self.bounds, self.depth = config, bounds, depth
        md = _get(config, "tree.max_depth", None)
        self.max_depth = int(md) if md is not None else 1      # configs without a tree section: one outer subdivision
        self.count = int(_get(config, "tree.subdivision_outer_count" if depth == 0 else "tree.subdivision_inner_count", 1) or 1)
        self.weight, self.sparse, self.children = 0.0, True, []

    def subdivide(self):
        if self.depth >= self.max_depth:
            return
        lo = self.bounds[0]
        extent = self.bounds[1] - lo
        n = self.count
        for i in range(n):
            for g in range(n):
                for h in range(n):
                    a = torch.tensor([i, g, h], dtype=torch.float) / n * extent          # same op order as the reference
                    b = torch.tensor([i + 1, g + 1, h + 1], dtype=torch.float) / n * extent
                    self.children.append(Node(self.config, (lo + a, lo + b), self.depth + 1))

    def clear(self):
        self.children = []


class TreeSampling:
    """src/nerf/tree.py:39-212.  `voxels` (V,2,3) and `memm` (V,) live on `device`; `root.children` is the flat leaf list."""

    def __init__(self, config, device=None):
        self.config = config
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.ray_near, self.ray_far = _get(config, "dataset.near"), _get(config, "dataset.far")
        self.ray_mean = (self.ray_near + self.ray_far) / 2
        bounds = torch.tensor([self.ray_near - self.ray_mean] * 3), torch.tensor([self.ray_far - self.ray_mean] * 3)
        self.root = Node(config, bounds, 0)
        self.root.subdivide()
        self.voxels, self.memm, self.counter = None, None, 1
        self.version = 0                       # bumped whenever `voxels` is replaced (BuFFModel re-uploads the list on change)
        self.engine = None                     # set by BuFFModel: the handle whose kernels run the integration
        self.consolidate()

    def set_voxels(self, voxels):
        """Adopt a flat (V,2,3) voxel list (e.g. exported from a checkpoint) without its node graph; weights restart."""
        self.voxels = torch.as_tensor(voxels).float().to(self.device)
        self.memm = torch.zeros(self.voxels.shape[0], device=self.device)
        self.counter = 1
        self.version += 1

    def touch(self):
        """Call after editing `voxels` in place."""
        self.version += 1

    # ------------------------------------------------------------------ schedule
    def ticked(self, step):
        off = int(_get(self.config, "tree.step_size_integration_offset", 0))
        every = int(_get(self.config, "tree.step_size_tree", 1))
        if step > off:
            cur = step - off
            return cur > 0 and cur % every == 0
        return False

    # ------------------------------------------------------------------ prune + subdivide (tree.py:127-175)
    def consolidate(self, split=False):
        if self.memm is not None:
            eps = float(_get(self.config, "tree.eps", 0.0))
            cap = int(_get(self.config, "tree.max_voxel_count", 1 << 30))
            inner = int(_get(self.config, "tree.subdivision_inner_count", 1)) ** 3 - 1
            memm = self.memm.detach().float().cpu()
            keep = memm > eps
            kept = [self.root.children[i] for i in torch.nonzero(keep).flatten().tolist()]
            inv_w = (1.0 - memm[keep]).tolist()
            # nodes closer to the root first, heavier first within a depth (python's sort is stable, like the reference's)
            order = sorted(range(len(kept)), key=lambda j: (kept[j].depth, inv_w[j]))
            n_kept, leaves = len(kept), []
            for pos, j in enumerate(order):
                node = kept[j]
                if len(leaves) + inner + n_kept - pos < cap:
                    node.subdivide()
                    leaves += node.children if len(node.children) > 0 else [node]
                else:
                    leaves.append(node)
            self.root.children = leaves
        if len(self.root.children) == 0:
            raise L.NmError(f"tree.eps = {_get(self.config, 'tree.eps')} pruned every voxel")
        self.voxels = torch.stack([torch.stack(tuple(n.bounds), 0) for n in self.root.children], 0).to(self.device)
        self.memm = torch.zeros(self.voxels.shape[0], device=self.device)
        self.counter = 1
        self.version = getattr(self, "version", 0) + 1

    # ------------------------------------------------------------------ per-step weight accumulation (tree.py:177-206)
    def ray_batch_integration(self, step, ray_voxel_indices, ray_batch_weights, ray_batch_weights_mask):
        """indices / weights / mask rows of the rays that hit a voxel (the reference passes x[mask]); rows may also be the
        full batch with index -1 on rays without a hit (what nm_ray_voxel_indices writes)."""
        if step < int(_get(self.config, "tree.step_size_integration_offset", 0)):
            return
        if self.engine is None or not ray_batch_weights.is_cuda:
            raise L.NmError("ray_batch_integration runs on the library's device kernels: CUDA tensors and a bound engine needed")
        if not self.memm.is_cuda:
            self.memm = self.memm.to(ray_batch_weights.device)
        self.engine.tree_integrate(ray_voxel_indices, ray_batch_weights, ray_batch_weights_mask, self.memm, self.counter)
        self.counter += 1

    # ------------------------------------------------------------------ wireframe for the tree logger (tree.py:104-125)
    _corner_axes = [[], [0], [1], [2], [0, 1], [1, 2], [0, 2], [0, 1, 2]]
    _faces = [0, 2, 1, 2, 4, 1, 0, 3, 2, 2, 3, 5, 0, 1, 6, 6, 3, 0, 1, 4, 7, 7, 6, 1, 3, 6, 7, 7, 5, 3, 2, 7, 4, 7, 2, 5]
    _colors = [[0, 0, 0], [128] * 3, [128] * 3, [128] * 3, [0, 0, 0], [128] * 3, [0, 0, 0], [128] * 3]

    def flatten(self):
        verts, faces, colors = [], [], []
        for node in self.root.children:
            extent = node.bounds[1] - node.bounds[0]
            base = len(verts)
            for axes in self._corner_axes:
                p = node.bounds[0].clone()
                p[axes] += extent[axes]
                verts.append(p)
            colors.append(torch.tensor(self._colors, dtype=torch.int))
            faces.append(torch.tensor(self._faces) + base)
        return torch.stack(verts, 0), torch.stack(faces, 0).view(-1, 3).int(), torch.stack(colors, 0).view(-1, 3)

    # ------------------------------------------------------------------ checkpoints (tree.py:345-358)
    def serialize(self):
        return {"root": self.root, "voxels": self.voxels, "memm": self.memm, "counter": self.counter}

    def deserialize(self, d):
        self.root = d["root"]
        self.voxels = d["voxels"].float().to(self.device)
        self.memm = d["memm"].float().to(self.device) if d.get("memm") is not None else torch.zeros(self.voxels.shape[0], device=self.device)
        self.counter = d["counter"]
        self.version = getattr(self, "version", 0) + 1
