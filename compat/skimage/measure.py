"""`skimage.measure.marching_cubes(volume, level)` seam (src/mesh_nerf.py:79) -> the CUDA marching cubes."""
from nerfmeshes_b200.mesh import marching_cubes  # noqa: F401
