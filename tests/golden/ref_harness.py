"""Import the UNMODIFIED reference (/root/reference/src) in this container.

Test-infrastructure only.  The reference needs a handful of third-party
packages that are not installed here (pytorch_lightning, pytorch3d, skimage,
imageio, OpenEXR, matplotlib); none of them is on the render / grid hot path, so
they are replaced by inert stub modules (recipe: SURVEY.md Appendix C).  This
module is used ONLY by `make_golden.py` to generate the fixtures committed under
tests/golden/ — /root/reference does not exist on the GPU box, so nothing in the
test-suite proper imports this file.
"""
import collections
import collections.abc
import sys
import types

import torch

REF_ROOT = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class AttributeDict(dict):
    """Unpickling target used inside the PL-0.9 checkpoints."""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class _LightningModule(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.global_step = 0
        self.trainer = None
        self.logger = None

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def load_from_checkpoint(cls, path, map_location="cpu"):
        ck = torch.load(path, map_location=map_location, weights_only=False)
        m = cls(dict(ck["hyper_parameters"]))
        if hasattr(m, "on_load_checkpoint"):
            m.on_load_checkpoint(ck)
        m.load_state_dict(ck["state_dict"], strict=False)
        return m


def install():
    """Install stubs + sys.path entry; idempotent."""
    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_nm_stub", False):
        return
    collections.MutableMapping = collections.abc.MutableMapping
    _mod("pytorch_lightning", LightningModule=_LightningModule, Trainer=object,
         seed_everything=torch.manual_seed, _nm_stub=True)
    _mod("pytorch_lightning.utilities")
    _mod("pytorch_lightning.utilities.parsing", AttributeDict=AttributeDict)
    _mod("pytorch_lightning.callbacks", Callback=object, ModelCheckpoint=object)
    _mod("pytorch_lightning.loggers",
         TensorBoardLogger=type("TensorBoardLogger", (), {"NAME_HPARAMS_FILE": "hparams.yaml"}))
    _mod("pytorch3d")
    _mod("pytorch3d.ops", sample_points_from_meshes=None)
    _mod("pytorch3d.loss", chamfer_distance=None)
    _mod("pytorch3d.structures", Meshes=None)
    _mod("skimage", measure=None)
    _mod("skimage.measure")
    for n in ("imageio", "OpenEXR", "Imath", "matplotlib", "matplotlib.pyplot"):
        _mod(n)
    sys.path.insert(0, REF_ROOT + "/src")


def ckpt_path(name):
    return f"{REF_ROOT}/pretrained/{name}/default/version_0/checkpoints/model_last.ckpt"


def load_model(kind, name):
    install()
    import models  # noqa: the reference's package
    cls = getattr(models, kind)
    return cls.load_from_checkpoint(ckpt_path(name)).eval()
