"""TensorBoardLogger with Lightning 0.9's directory scheme (save_dir/name/version_<n>) over torch.utils.tensorboard: the
reference's PathParser builds it, LoggerCallback calls log_metrics, the validation steps call experiment.add_image, and
eval_nerf.py / mesh_nerf.py later read <log_dir>/hparams.yaml (NAME_HPARAMS_FILE) to rebuild the config."""
import os

import yaml


class TensorBoardLogger:
    NAME_HPARAMS_FILE = "hparams.yaml"

    def __init__(self, save_dir=None, name="default", version=None, **unused):
        self.save_dir, self.name = str(save_dir) if save_dir is not None else None, name
        self._version = version
        self._experiment = None
        self.hparams = {}

    @property
    def root_dir(self):
        return os.path.join(self.save_dir, self.name) if self.name else self.save_dir

    @property
    def version(self):
        if self._version is None:
            root = self.root_dir
            existing = []
            if os.path.isdir(root):
                for d in os.listdir(root):
                    if d.startswith("version_") and d[8:].isdigit():
                        existing.append(int(d[8:]))
            self._version = max(existing) + 1 if existing else 0
        return self._version

    @property
    def log_dir(self):
        v = self.version
        return os.path.join(self.root_dir, v if isinstance(v, str) else f"version_{v}")

    @property
    def experiment(self):
        if self._experiment is None:
            from torch.utils.tensorboard import SummaryWriter
            os.makedirs(self.log_dir, exist_ok=True)
            self._experiment = SummaryWriter(log_dir=self.log_dir)
        return self._experiment

    def log_metrics(self, metrics, step=None):
        for k, v in metrics.items():
            self.experiment.add_scalar(k, float(v), step)

    def log_hyperparams(self, params):
        self.hparams = dict(params)
        os.makedirs(self.log_dir, exist_ok=True)
        path = os.path.join(self.log_dir, self.NAME_HPARAMS_FILE)
        if not os.path.exists(path):
            plain = {k: (v if isinstance(v, (int, float, str, bool, list, type(None))) else str(v)) for k, v in self.hparams.items()}
            with open(path, "w") as f:
                yaml.dump(plain, f)

    def save(self):
        if self._experiment is not None:
            self._experiment.flush()

    def finalize(self, status):
        self.save()
