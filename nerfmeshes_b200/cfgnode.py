"""Attribute-access config container + the flat/nested conversions the reference's models use.

Mirrors the behaviour (not the code) of src/nerf/cfgnode.py:36-142 (`CfgNode(dict)` with attribute access) and
src/models/model_helpers.py:6-29 (`flatten_dict` / `nest_dict` with a separator).
"""
from collections.abc import Mapping


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, Mapping) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def nest_dict(flat, sep="."):
    """{'a.b': 1} -> {'a': {'b': 1}}; already-nested mappings pass through."""
    out = {}
    for key, value in flat.items():
        if isinstance(value, Mapping):
            value = nest_dict(value, sep)
        node = out
        parts = str(key).split(sep)
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        if isinstance(value, dict) and isinstance(node.get(parts[-1]), dict):
            node[parts[-1]].update(value)
        else:
            node[parts[-1]] = value
    return out


def flatten_dict(d, parent_key="", sep="."):
    items = {}
    for k, v in d.items():
        key = f"{parent_key}{sep}{k}" if parent_key else str(k)
        if isinstance(v, Mapping):
            items.update(flatten_dict(v, key, sep))
        else:
            items[key] = v
    return items
