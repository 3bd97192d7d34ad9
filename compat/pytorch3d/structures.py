class Meshes:
    """Placeholder: only mesh_nerf.create_mesh (chamfer evaluation, off in every shipped config) would use it."""
    def __init__(self, verts=None, faces=None):
        self.verts, self.faces = verts, faces
