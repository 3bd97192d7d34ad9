"""A minimal PyTorch-Lightning 0.9 surface for the reference's scripts on the fused path: `LightningModule` (a
torch.nn.Module; the hooks live in nerfmeshes_b200.lightning) and a `Trainer` whose `fit(model)` runs the loop train_nerf.py
expects (src/train_nerf.py:76-101): setup -> configure_optimizers -> [resume] -> epochs of training_step / optimiser step /
scheduler step -> validation every `check_val_every_n_epoch` -> ModelCheckpoint -> callbacks with Lightning's hook names and
arguments (the reference's LoggerCallback runs unmodified).  One process, one GPU: the multi-GPU data-parallel step is
nerfmeshes_b200.train.training_step + parallel.allreduce_gradients under torchrun, not Lightning's DDP plugin."""
import os

import torch

from .callbacks import Callback, ModelCheckpoint  # noqa: F401


class LightningModule(torch.nn.Module):
    pass


def seed_everything(seed):
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    return seed


class Trainer:
    def __init__(self, logger=None, checkpoint_callback=None, callbacks=None, resume_from_checkpoint=None, gpus=1,
                 default_root_dir=None, max_steps=None, max_epochs=1000, deterministic=False, num_sanity_val_steps=0,
                 accumulate_grad_batches=1, precision=32, **unused):
        if precision != 32:
            raise ValueError("the fused path computes in fp16x3-split / fp32-accumulate; --precision 16 has no meaning here")
        if accumulate_grad_batches != 1:
            raise NotImplementedError("accumulate_grad_batches != 1")
        self.logger, self.checkpoint_callback = logger, checkpoint_callback
        self.callbacks = list(callbacks or [])
        self.resume_from_checkpoint, self.gpus, self.default_root_dir = resume_from_checkpoint, gpus, default_root_dir
        self.max_steps, self.min_steps, self.max_epochs, self.min_epochs = max_steps, None, max_epochs, 1
        self.check_val_every_n_epoch = 1
        self.global_step, self.current_epoch, self.batch_idx = 0, 0, 0
        self.optimizers, self.lr_schedulers = [], []
        self.callback_metrics = {}
        self.train_dataloader, self.val_dataloaders = None, None
        self.model = None
        if deterministic:
            torch.backends.cudnn.deterministic = True

    # ------------------------------------------------------------------ helpers
    def _call(self, hook, *args):
        for cb in self.callbacks + ([self.checkpoint_callback] if self.checkpoint_callback is not None else []):
            fn = getattr(cb, hook, None)
            if fn is not None:
                fn(self, self.model, *args)

    def _metrics(self, log):
        for k, v in (log or {}).items():
            self.callback_metrics[k] = float(v)

    def save_checkpoint(self, path):
        sched = self.lr_schedulers[0]["scheduler"] if self.lr_schedulers else None
        ck = self.model.save_checkpoint(path, global_step=self.global_step, optimizer=self.optimizers[0] if self.optimizers else None,
                                        lr_scheduler=sched)
        return ck

    def _restore(self, model):
        path = self.resume_from_checkpoint
        if not path or not os.path.exists(str(path)):
            return
        from nerfmeshes_b200.models import load_lightning_checkpoint
        ck = load_lightning_checkpoint(str(path))
        if hasattr(model, "on_load_checkpoint"):
            model.on_load_checkpoint(ck)
        model.load_state_dict(ck["state_dict"], strict=False)
        self.global_step = int(ck.get("global_step", 0))
        self.current_epoch = int(ck.get("epoch", 0))
        for opt, st in zip(self.optimizers, ck.get("optimizer_states", [])):
            opt.load_state_dict(st)
        for sch, st in zip(self.lr_schedulers, ck.get("lr_schedulers", [])):
            sch["scheduler"].load_state_dict(st)
        print(f"Restored {path} at step {self.global_step}")

    # ------------------------------------------------------------------ the loop
    def run_validation(self):
        model = self.model
        was_training = model.training
        model.eval()
        self._call("on_validation_start")
        outputs = []
        with torch.no_grad():
            for i, batch in enumerate(self.val_dataloaders):
                outputs.append(model.validation_step(batch, i))
                self._call("on_validation_batch_end", batch, i, 0)
        if outputs:
            res = model.validation_epoch_end(outputs)
            self._metrics(res.get("log"))
            self.callback_metrics["val_loss"] = float(res["val_loss"])
            if self.logger is not None:
                self.logger.log_metrics({k: float(v) for k, v in res.get("log", {}).items()}, step=self.global_step)
        self._call("on_validation_epoch_end")
        self._call("on_validation_end")
        model.train(was_training)

    def fit(self, model):
        self.model = model
        model.trainer, model.logger = self, self.logger
        if self.gpus and torch.cuda.is_available():
            model.cuda(int(os.environ.get("LOCAL_RANK", "0")))
        # (without a GPU nothing is moved and the first compute call — ray generation in the dataset, or the first step —
        #  raises the library's "needs a CUDA device" error: there is no CPU path to fall back to)
        if hasattr(model, "setup"):
            model.setup("fit")
        optimizers, schedulers = model.configure_optimizers()
        self.optimizers, self.lr_schedulers = list(optimizers), list(schedulers)
        self._restore(model)
        model.global_step = self.global_step
        if self.logger is not None and hasattr(self.logger, "log_hyperparams"):
            self.logger.log_hyperparams(getattr(model, "hparams", {}))
        self.train_dataloader = model.train_dataloader()
        self.val_dataloaders = model.val_dataloader()
        every = max(int(self.check_val_every_n_epoch or 1), 1)
        model.train()
        self._call("on_fit_start")
        self._call("on_train_start")
        done = False
        for epoch in range(self.current_epoch, max(int(self.max_epochs or 1), 1)):
            self.current_epoch = epoch
            self._call("on_train_epoch_start")
            for batch_idx, batch in enumerate(self.train_dataloader):
                self.batch_idx = batch_idx
                out = model.training_step(batch, batch_idx)
                loss = out["loss"] if isinstance(out, dict) else out
                if isinstance(loss, torch.Tensor) and loss.requires_grad:     # a step that left the backward to the trainer
                    for opt in self.optimizers:
                        opt.zero_grad()
                    loss.backward()
                for opt in self.optimizers:
                    opt.step()
                for sch in self.lr_schedulers:
                    if sch.get("interval", "epoch") == "step":
                        sch["scheduler"].step()
                self.global_step += 1
                model.global_step = self.global_step
                if isinstance(out, dict):
                    self._metrics(out.get("log"))
                self._call("on_train_batch_end", batch, batch_idx, 0)
                if self.max_steps is not None and self.global_step >= self.max_steps:
                    done = True
                    break
            for sch in self.lr_schedulers:
                if sch.get("interval", "epoch") == "epoch":
                    sch["scheduler"].step()
            self._call("on_train_epoch_end")
            if (epoch + 1) % every == 0 or done:
                self.run_validation()
            if done:
                break
        self._call("on_train_end")
        self._call("on_fit_end")
        if self.logger is not None and hasattr(self.logger, "finalize"):
            self.logger.finalize("success")
        return 1
