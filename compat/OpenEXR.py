"""Inert placeholder (EXR depth maps are read only by the dataset loaders)."""
