"""Seq2Seq container: the model-level API boundary (contract: reference models/seq2seq.py:7-140)."""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from ..config import Constants
from ..engine import Engine


class VocabProjection(nn.Linear):
    """``model.tgt_word_prj``: nn.Linear parameters, forward = navc GEMM (callers such as
    decoding/algorithms.py:149 of the reference invoke the attribute directly)."""

    def forward(self, hidden):
        owner = self._owner() if getattr(self, "_owner", None) else None
        if owner is None:
            raise RuntimeError("tgt_word_prj must be built through models.get_model")
        if torch.is_grad_enabled() and owner.training:
            from ..training import vocab_forward_train
            return vocab_forward_train(owner.engine, self, hidden)
        with torch.no_grad():
            shape = hidden.shape
            out = owner.engine.logits(hidden.reshape(-1, shape[-1]))
            return out.view(*shape[:-1], -1)


class Seq2Seq(nn.Module):
    def __init__(self, opt, preEncoder=None, encoder=None, joint_representation_learner=None,
                 auxiliary_task_predictor=None, decoder=None, tgt_word_prj=None, **kwargs):
        super().__init__()
        self.opt = opt
        self.preEncoder = preEncoder
        self.encoder = encoder
        self.joint_representation_learner = joint_representation_learner
        self.auxiliary_task_predictor = auxiliary_task_predictor
        self.decoder = decoder
        self.tgt_word_prj = tgt_word_prj
        if opt.get("tie_weights", False):
            self._tie_weights(opt["vocab_size"])
        ref = weakref.ref(self)
        object.__setattr__(self.tgt_word_prj, "_owner", ref)
        inner = getattr(self.decoder, "bert", self.decoder)
        object.__setattr__(inner, "_engine_ref", ref)
        self.__dict__["engine"] = Engine(self)

    def _tie_weights(self, vocab_size):
        self.tgt_word_prj.weight = self.decoder.get_word_embeddings().weight
        self.tgt_word_prj.bias = nn.Parameter(torch.zeros(vocab_size).float(), requires_grad=True)

    def set_precision(self, precision):
        sink = getattr(self.__dict__.get("engine"), "grad_sink", None)
        self.__dict__["engine"] = Engine(self, precision)
        self.engine.grad_sink = sink  # a GradientAllReduce attached earlier keeps accumulating in place
        return self

    # -- reference seq2seq.py:35-63 -------------------------------------------------------------
    def encode(self, feats, **kwargs):
        if self.opt.get("automatic_mask", False):
            raise NotImplementedError("automatic_mask is not used by the method presets")
        if torch.is_grad_enabled() and self.training:
            from ..training import encode_train
            return encode_train(self, feats)
        with torch.no_grad():
            results = self.engine.encode(list(feats))
        # let a later model.decoder(enc_output=...) call reuse the bf16 copies / frame mean
        try:
            results["enc_output"]._navc_cache = results["_navc"]
        except Exception:
            pass
        return results

    # -- reference seq2seq.py:65-80 --------------------------------------------------------------
    def prepare_inputs_for_decoder(self, encoder_outputs, category):
        inputs = {"category": category, "enc_output": encoder_outputs["enc_output"]}
        if isinstance(inputs["enc_output"], list):
            assert len(inputs["enc_output"]) == 1
            inputs["enc_output"] = inputs["enc_output"][0]
        return inputs

    def forward(self, **kwargs):
        fn = getattr(self, "forward_" + self.opt["decoding_type"], None)
        if fn is None:
            raise ValueError("unsupported decoding_type %r" % self.opt["decoding_type"])
        return fn(kwargs)

    def _plan_train(self, feats, tgt_tokens):
        """Training only: decide the packed-row layout of this batch before anything is launched (training.py)."""
        if torch.is_grad_enabled() and self.training and tgt_tokens is not None:
            from ..training import plan_packing_ahead
            feats = list(feats)
            plan_packing_ahead(self, tgt_tokens if isinstance(tgt_tokens, (list, tuple)) else [tgt_tokens],
                               sum(int(f.shape[1]) for f in feats))

    def _decode_and_project(self, results, tgt_tokens, category, **dec_kwargs):
        inputs = self.prepare_inputs_for_decoder(results, category)
        hidden_states, embs, *_ = self.decoder(tgt_seq=tgt_tokens, **inputs, **dec_kwargs)
        if not isinstance(hidden_states, list):
            hidden_states = [hidden_states]
        logprobs = []
        for h in hidden_states:
            if torch.is_grad_enabled() and self.training:
                from ..training import LazyLogProbs, vocab_logprobs_train
                if self.opt.get("navc_fused_ce", False):
                    # consumed by navc_b200.misc.crit: projection + log-softmax + masked NLL fused, no [B,S,V] tensor
                    logprobs.append(LazyLogProbs(self, h))
                else:
                    logprobs.append(vocab_logprobs_train(self, h))  # projection + log-softmax, one autograd node
                continue
            logits = self.tgt_word_prj(h)
            shape = logits.shape
            logprobs.append(self.engine.log_softmax_(logits.view(-1, shape[-1])).view(shape))
        results[Constants.mapping["lang"][0]] = logprobs
        results.pop("_navc", None)
        return results

    # -- reference seq2seq.py:86-108 -------------------------------------------------------------
    def forward_NARFormer(self, kwargs):
        feats, tgt_tokens, category = (kwargs.get(k, None) for k in ("feats", "tgt_tokens", "category"))
        self._plan_train(feats, tgt_tokens)
        results = self.encode(feats)
        return self._decode_and_project(results, tgt_tokens, category)

    # -- reference seq2seq.py:110-140 ------------------------------------------------------------
    def forward_ARFormer(self, kwargs):
        feats, tgt_tokens, category = (kwargs.get(k, None) for k in ("feats", "tgt_tokens", "category"))
        decoding_type = kwargs.get("decoding_type", self.opt["decoding_type"])
        cut = (lambda t: t[:, 1:]) if decoding_type == "SelfMask" else (lambda t: t[:, :-1])
        tgt_tokens = [cut(t) for t in tgt_tokens] if isinstance(tgt_tokens, list) else cut(tgt_tokens)
        self._plan_train(feats, tgt_tokens)
        results = self.encode(feats)
        return self._decode_and_project(results, tgt_tokens, category, decoding_type=decoding_type,
                                        output_attentions=kwargs.get("output_attentions", False))
