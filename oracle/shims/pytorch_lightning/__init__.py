"""Import plumbing, not a product feature: the reference imports `pytorch_lightning` (predict_mos.py:4,
models/models.py:14, dataloader/datasets.py:11) which is not installed in this image.  Only what the
inference path touches is provided: LightningModule.{save_hyperparameters, hparams, load_from_checkpoint, log}."""
from .core.lightning import LightningModule

__version__ = "1.5.10+insmos_b200-shim"


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Trainer:
    def __init__(self, *a, **k):
        raise NotImplementedError("pytorch_lightning.Trainer: training loop is out of scope (SURVEY 8f N3)")
