"""`models` namespace of the reference (src/models/__init__.py:1-4), backed by nerfmeshes_b200."""
from nerfmeshes_b200.models import BaseModel, BuFFModel, NeRFModel  # noqa: F401
from . import model_helpers  # noqa: F401
from .model_helpers import *  # noqa: F401,F403
