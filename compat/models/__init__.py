"""`models` namespace of the reference (src/models/__init__.py:1-4), backed by nerfmeshes_b200: the engine-backed models plus
the Lightning hooks train_nerf.py's Trainer drives (nerfmeshes_b200.lightning)."""
import pytorch_lightning as pl
from nerfmeshes_b200 import models as _m
from nerfmeshes_b200.lightning import LightningHooks


class BaseModel(LightningHooks, _m.BaseModel, pl.LightningModule):
    pass


class NeRFModel(LightningHooks, _m.NeRFModel, pl.LightningModule):
    pass


class BuFFModel(LightningHooks, _m.BuFFModel, pl.LightningModule):
    pass


from . import model_helpers  # noqa: E402,F401
from .model_helpers import *  # noqa: E402,F401,F403
