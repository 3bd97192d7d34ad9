"""`nerf.tree` of the reference (src/nerf/tree.py): the host-side voxel tree; its per-step kernels live in the library."""
from nerfmeshes_b200.tree import Node, TreeSampling  # noqa: F401
