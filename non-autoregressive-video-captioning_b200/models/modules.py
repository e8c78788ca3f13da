"""Parameter containers with the reference's module tree (=> identical ``state_dict`` keys and
shapes) whose arithmetic is executed by the navc kernels, not by these
modules.  Construction order follows the reference so that ``torch.manual_seed(s); get_model(opt)``
draws the same initial weights (reference: models/Encoder.py:9-66, joint_representation.py:5-22,
Predictor.py:12-43, bert.py:46-260, Decoder.py:67-88, 181-186).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..config import Constants


class _Holder(nn.Module):
    """A module that only owns parameters; calling it directly is a bug in the host code."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("%s holds parameters only; compute runs in libnavc" % type(self).__name__)


class HighWay(_Holder):
    def __init__(self, hidden_size, with_gate=True):
        super().__init__()
        self.with_gate = with_gate
        self.w1 = nn.Linear(hidden_size, hidden_size)
        if with_gate:
            self.w2 = nn.Linear(hidden_size, hidden_size)
        self.tanh = nn.Tanh()


class Encoder_HighWay(_Holder):
    def __init__(self, opt):
        super().__init__()
        with_gate = opt.get("gate", True)
        self.num_feats = len(opt["modality"])
        for ch in opt["modality"].lower():
            in_dim = opt.get("dim_" + ch, None)
            assert in_dim is not None, "modality %s needs dim_%s in opt" % (opt["modality"], ch)
            out_dim = opt.get("dim_hidden", 512)
            self.add_module("Encoder_%s" % ch.upper(), nn.Sequential(
                nn.Linear(in_dim, out_dim), HighWay(out_dim, with_gate), nn.Dropout(opt.get("encoder_dropout", 0.5))))


class Joint_Representaion_Learner(_Holder):
    def __init__(self, feats_size, opt):
        super().__init__()
        self.fusion = opt.get("fusion", "temporal_concat")
        if self.fusion not in ("temporal_concat", "addition", "none"):
            raise ValueError("We now only support the fusion type: temporal_concat | addition | none")
        self.is_bn = opt.get("norm_type", "bn").lower() == "bn"
        if not opt["no_encoder_bn"]:
            if self.fusion == "addition":
                feats_size = [feats_size[0]]
            for i, d in enumerate(feats_size):
                self.add_module("%s%d" % ("bn" if self.is_bn else "ln", i), nn.BatchNorm1d(d) if self.is_bn else nn.LayerNorm(d))


class BertEmbeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg["vocab_size"], cfg["dim_hidden"], padding_idx=Constants.PAD)
        self.position_embeddings = nn.Embedding(cfg["max_len"], cfg["dim_hidden"])
        self.category_embeddings = nn.Embedding(cfg["num_category"], cfg["dim_hidden"]) if cfg["with_category"] else None
        self.LayerNorm = nn.LayerNorm(cfg["dim_hidden"], eps=cfg["layer_norm_eps"])
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])


class BertSelfAttention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        d = cfg["dim_hidden"]
        if d % cfg["num_attention_heads"] != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (d, cfg["num_attention_heads"]))
        self.query, self.key, self.value = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
        self.dropout = nn.Dropout(cfg["attention_probs_dropout_prob"])


class BertSelfOutput(_Holder):
    def __init__(self, cfg):
        super().__init__()
        d = cfg["dim_hidden"]
        self.dense = nn.Linear(d, d)
        self.LayerNorm = nn.LayerNorm(d, eps=cfg["layer_norm_eps"]) if cfg["with_layernorm"] else None
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])


class BertAttention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.self = BertSelfAttention(cfg)
        self.output = BertSelfOutput(cfg)


class BertIntermediate(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg["dim_hidden"], cfg["intermediate_size"])


class BertOutput(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg["intermediate_size"], cfg["dim_hidden"])
        self.LayerNorm = nn.LayerNorm(cfg["dim_hidden"], eps=cfg["layer_norm_eps"]) if cfg["with_layernorm"] else None
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])


class BertLayer(_Holder):
    def __init__(self, cfg):
        super().__init__()
        if cfg.get("pos_attention", False):
            raise NotImplementedError("pos_attention is not used by any method preset (config/methods.yaml)")
        self.attention = BertAttention(cfg)
        self.pos_attention = None
        self.attend_to_enc_output = BertAttention(cfg)
        self.intermediate = BertIntermediate(cfg)
        self.output = BertOutput(cfg)
