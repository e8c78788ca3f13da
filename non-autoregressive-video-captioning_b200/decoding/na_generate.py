"""Non-autoregressive generation entry point (contract: reference decoding/na_generate.py:14-108).

``generate(opt, model, teacher_model, encoder_outputs, teacher_encoder_outputs, category, tgt_tokens,
tgt_vocab, dict_mapping, length_bias, **kwargs) -> (hypotheses [B, Smax] int64, None)``.

Differences from the reference that do not change results: encoder memory is NOT repeated
x length_beam_size (kernels index ``row // lbs``); cross-attention K|V are projected once per video
(SURVEY.md F6); no logits/probability tensors are materialised.
"""
from __future__ import annotations

import torch

from .. import _lib as L
from ..config import Constants
from .algorithms import ALGORITHMS, Refiner

_UNSUPPORTED = ("load_generated_captions", "collect_best_candidate_iterative_results", "collect_last", "example",
                "manual")


def generate(opt, model, teacher_model, encoder_outputs, teacher_encoder_outputs, category, tgt_tokens,
             tgt_vocab, dict_mapping, length_bias, **kwargs):
    paradigm = opt.get("paradigm", "mp")
    assert paradigm in ALGORITHMS, paradigm
    for key in _UNSUPPORTED:
        if opt.get(key, False):
            raise NotImplementedError("opt[%r] (analysis-only path of the reference) is not supported" % key)
    if kwargs.get("output_attentions", False):
        raise NotImplementedError("output_attentions during generation (only consumed with opt['example'])")
    eng = model.engine
    eng.sync_weights()
    pred_length = encoder_outputs["pred_length"].contiguous()
    B, max_len = pred_length.shape
    lbs = int(opt["length_beam_size"])
    dev = pred_length.device

    # length beam (na_generate.py:33-37, 116-135) + the one host read of the batch-wide Smax
    beam = torch.empty((B, lbs), dtype=torch.int32, device=dev)
    smax = torch.zeros((1,), dtype=torch.int32, device=dev)
    L.call("navc_length_beam", L.ptr(pred_length), B, max_len, lbs, int(length_bias), L.ptr(beam), L.ptr(smax), L.stream())
    S = int(smax.item())

    mem = eng.memory(encoder_outputs["enc_output"], encoder_outputs.get("_navc"))
    tmem = None
    if teacher_model is not None and teacher_encoder_outputs is not None:
        teacher_model.engine.sync_weights()
        tmem = teacher_model.engine.memory(teacher_encoder_outputs["enc_output"], teacher_encoder_outputs.get("_navc"))
    else:
        teacher_model = None
    cat = category.contiguous() if category is not None else None

    ref = Refiner(opt, model, teacher_model, mem, tmem, cat, beam, S, dict_mapping)
    tokens, lprobs = ALGORITHMS[paradigm](ref)

    hyp = torch.empty((B, S), dtype=torch.int64, device=dev)
    L.call("navc_select_best", L.ptr(tokens), L.ptr(lprobs), L.ptr(ref.lens), B, lbs, S,
           float(opt.get("beam_alpha", 1.0)), L.ptr(hyp), None, L.stream())
    generate.last_stats = {"passes": ref.passes, "S": S, "N": ref.N, "steps": ref.n_steps}
    return hyp, None


generate.last_stats = {}
