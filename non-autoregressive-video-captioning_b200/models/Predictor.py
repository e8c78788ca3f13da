"""Auxiliary predictors (contract: reference models/Predictor.py:12-43).  Discovered reflectively by
the ``Predictor_`` name prefix, as models/__init__.py:41-52 of the reference does."""
import torch.nn as nn

from .modules import _Holder

__all__ = ("Predictor_length", "Auxiliary_Task_Predictor")


class Predictor_length(_Holder):
    """mean_t(enc_output) -> Linear -> ReLU -> Dropout -> Linear(max_len) -> log_softmax; the
    arithmetic is the `navc_length_head` kernel (engine.Engine.encode)."""

    def __init__(self, opt, key_name):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(opt["dim_hidden"], opt["dim_hidden"]), nn.ReLU(),
            nn.Dropout(opt["hidden_dropout_prob"]), nn.Linear(opt["dim_hidden"], opt["max_len"]))
        self.key_name = key_name


class Auxiliary_Task_Predictor(_Holder):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)
