"""The two helpers of src/models/model_helpers.py that other modules import (cfg nesting, o + t*d)."""
from nerfmeshes_b200.cfgnode import flatten_dict as _flatten, nest_dict as _nest


def nest_dict(flat, sep="_"):
    return _nest(flat, sep)


def flatten_dict(d, parent_key="", sep="_"):
    return _flatten(d, parent_key, sep)


def intervals_to_ray_points(point_intervals, ray_directions, ray_origin):
    """src/models/model_helpers.py:32-35 (host tensors; the fused kernel evaluates this per sample on the device)."""
    return ray_origin[..., None, :] + ray_directions[..., None, :] * point_intervals[..., :, None]
