import torch
import torch.nn as nn


class _HParams(dict):
    __getattr__ = dict.__getitem__


class LightningModule(nn.Module):
    def save_hyperparameters(self, hparams=None, *args, **kwargs):
        object.__setattr__(self, "_hparams", _HParams(hparams or {}))

    @property
    def hparams(self):
        return self._hparams

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, hparams=None, strict=True, **kwargs):
        """Lightning checkpoint: {'state_dict': {...'model.' prefixed keys...}, 'hyper_parameters': {...}}"""
        # InsMOS checkpoints hold a state_dict and a plain hyper_parameters dict: no arbitrary unpickling is needed.
        # Full pickle loading (code execution from an untrusted .ckpt) only behind INSMOS_UNSAFE_CKPT=1.
        import os
        unsafe = os.environ.get("INSMOS_UNSAFE_CKPT", "0") == "1"
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=not unsafe)
        hp = hparams if hparams is not None else ckpt.get("hyper_parameters")
        model = cls(hp, **kwargs)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model
