"""`nerf.models` (src/nerf/models.py): the one architecture every shipped config selects."""
from nerfmeshes_b200.models import FlexibleNeRFModel  # noqa: F401
