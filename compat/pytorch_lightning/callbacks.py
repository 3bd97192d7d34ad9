"""Callback base + ModelCheckpoint with the Lightning-0.9 constructor the reference uses (src/train_nerf.py:65-66):
ModelCheckpoint(filepath=<dir>, save_top_k=3, save_last=True, monitor="val_loss", mode="min", prefix="model_").  After every
validation: `<prefix>last.ckpt` (what --log-checkpoint resumes from, lightning_modules.py:PathParser) and the save_top_k best
`<prefix>epoch=<e>.ckpt` by the monitored metric."""
import os


class Callback:
    pass


class ModelCheckpoint(Callback):
    def __init__(self, filepath=None, monitor="val_loss", verbose=False, save_last=False, save_top_k=1, mode="min", prefix="", **unused):
        self.dirpath = str(filepath) if filepath is not None else None
        self.monitor, self.verbose, self.save_last, self.save_top_k, self.mode, self.prefix = monitor, verbose, save_last, save_top_k, mode, prefix
        self.best = []            # (score, path)
        self.best_model_path, self.best_model_score = "", None

    def on_validation_end(self, trainer, pl_module):
        if self.dirpath is None:
            return
        os.makedirs(self.dirpath, exist_ok=True)
        if self.save_last:
            trainer.save_checkpoint(os.path.join(self.dirpath, f"{self.prefix}last.ckpt"))
        score = trainer.callback_metrics.get(self.monitor)
        if score is None or not self.save_top_k:
            return
        key = score if self.mode == "min" else -score
        path = os.path.join(self.dirpath, f"{self.prefix}epoch={trainer.current_epoch}.ckpt")
        if self.save_top_k < 0 or len(self.best) < self.save_top_k or key < max(k for k, _ in self.best):
            trainer.save_checkpoint(path)
            self.best = sorted([(k, p) for k, p in self.best if p != path] + [(key, path)])
            while self.save_top_k > 0 and len(self.best) > self.save_top_k:
                _, drop = self.best.pop()
                if os.path.exists(drop):
                    os.remove(drop)
            self.best_model_score, self.best_model_path = self.best[0]
            if self.verbose:
                print(f"Epoch {trainer.current_epoch}: {self.monitor} = {score:.6f}, saved {path}")

    on_fit_end = on_validation_end
