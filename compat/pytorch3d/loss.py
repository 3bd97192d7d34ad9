def chamfer_distance(*a, **k):
    raise NotImplementedError("pytorch3d is not available; chamfer evaluation is out of scope")
