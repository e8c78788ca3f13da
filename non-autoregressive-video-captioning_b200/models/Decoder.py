"""BertDecoder / BertDecoderDisentangled with the reference call contract
(reference models/Decoder.py:67-215), executed by the navc kernels through ``engine.Engine``."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..config import Constants
from .modules import BertEmbeddings, BertLayer

__all__ = ("BertDecoder", "BertDecoderDisentangled")


class BertDecoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        self.embedding = BertEmbeddings(cfg)
        self.layer = nn.ModuleList([BertLayer(cfg) for _ in range(cfg["num_hidden_layers_decoder"])])
        self.pos_attention = cfg["pos_attention"]
        self.enhance_input = cfg["enhance_input"]
        self.watch = cfg["watch"]
        self.decoding_type = cfg["decoding_type"]
        self._engine_ref = None  # set by Seq2Seq (weak, avoids registering the model as a submodule)

    def get_word_embeddings(self):
        return self.embedding.word_embeddings

    def set_word_embeddings(self, we):
        self.embedding.word_embeddings = we

    def _engine(self):
        if self._engine_ref is None or self._engine_ref() is None:
            raise RuntimeError("BertDecoder must be built through models.get_model (needs its Seq2Seq engine)")
        return self._engine_ref().engine

    def forward(self, tgt_seq, enc_output=None, category=None, signals=None, tags=None, **kwargs):
        """Returns ([hidden [N,S,D]], embs [N,D][, attentions]) like reference Decoder.py:96-178."""
        if signals is not None or tags is not None:
            raise NotImplementedError("signals/tags inputs are unused by the method presets")
        decoding_type = kwargs.get("decoding_type", self.decoding_type)
        output_attentions = kwargs.get("output_attentions", False)
        if isinstance(enc_output, list):
            assert len(enc_output) == 1
            enc_output = enc_output[0]
        eng = self._engine()
        if torch.is_grad_enabled() and self.training:
            from ..training import decoder_forward_train
            return decoder_forward_train(self, eng, tgt_seq, enc_output, category, decoding_type)
        with torch.no_grad():
            tgt_seq = tgt_seq.contiguous()
            mem = eng.memory(enc_output.contiguous().float(), getattr(enc_output, "_navc_cache", None))
            cat = category.contiguous() if category is not None else None
            hid, attns = eng.decoder_pass(tgt_seq, mem, 1, cat, decoding_type, want_attn=output_attentions, want_f32=True)
            N, S = tgt_seq.shape
            hidden = hid.f32.view(N, S, -1)
            non_pad = tgt_seq.ne(Constants.PAD).float().unsqueeze(-1)
            embs = hidden.sum(1) / non_pad.sum(1)  # returned, unused downstream (bert.py:301)
        outputs = ([hidden], embs)
        if output_attentions:
            outputs = outputs + (tuple(attns),)
        return outputs


class BertDecoderDisentangled(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.bert = BertDecoder(config)

    def get_word_embeddings(self):
        return self.bert.get_word_embeddings()

    def set_word_embeddings(self, we):
        self.bert.set_word_embeddings(we)

    def forward_(self, tgt_seq, enc_output, category, **kwargs):
        seq, embs, *rest = self.bert(tgt_seq, enc_output, category, **kwargs)
        if rest:
            return seq[0], embs, rest
        return seq[0], embs

    def forward(self, tgt_seq, enc_output, category, **kwargs):
        if isinstance(enc_output, list):
            assert len(enc_output) == 1
            enc_output = enc_output[0]
        if isinstance(tgt_seq, list):
            assert len(tgt_seq) == 2
            h1, _ = self.forward_(tgt_seq[0], enc_output, category, **kwargs)[:2]
            h2, embs = self.forward_(tgt_seq[1], enc_output, category, **kwargs)[:2]
            return ([h1, h2], embs)
        return self.forward_(tgt_seq, enc_output, category, **kwargs)
