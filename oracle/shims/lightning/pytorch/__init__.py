"""Minimal ``lightning.pytorch`` surface that reference ``pl_module.py`` / ``callbacks.py`` touch."""
import types

from torch import nn


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.logged = {}
        self.trainer = types.SimpleNamespace(estimated_stepping_batches=100)

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    def log_dict(self, d, **kwargs):
        self.logged.update(d)

    def on_validation_epoch_end(self):
        return None

    def on_test_epoch_end(self):
        return None


class Callback:
    pass


class Trainer:
    pass


def seed_everything(seed, workers=False):
    import random

    import numpy as np
    import torch

    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    return seed
