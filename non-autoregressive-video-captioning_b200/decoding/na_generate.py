"""Non-autoregressive generation entry point (contract: reference decoding/na_generate.py:14-108).

``generate(opt, model, teacher_model, encoder_outputs, teacher_encoder_outputs, category, tgt_tokens,
tgt_vocab, dict_mapping, length_bias, **kwargs) -> (hypotheses [B, Smax] int64, None)``.

Differences from the reference that do not change results: encoder memory is NOT repeated
x length_beam_size (kernels index ``row // lbs``); cross-attention K|V are projected once per video
(SURVEY.md F6); no logits/probability tensors are materialised.
"""
from __future__ import annotations

import os

import torch

from .. import _lib as L
from ..config import Constants
from .algorithms import ALGORITHMS, Refiner

_UNSUPPORTED = ("load_generated_captions", "collect_best_candidate_iterative_results", "collect_last", "example",
                "manual")


def graphs_enabled(opt) -> bool:
    """CUDA-graph replay of the refinement loop: opt['navc_graphs'] or $NAVC_GRAPHS (default on)."""
    v = opt.get("navc_graphs", None)
    if v is None:
        v = os.environ.get("NAVC_GRAPHS", "1")
    return str(v).lower() not in ("0", "false", "no", "off")


def _copy_act(dst, src):
    for name in ("f32", "hi", "lo"):
        d, s_ = getattr(dst, name), getattr(src, name)
        if d is not None:
            d.copy_(s_.view_as(d))


def _clone_inputs(mem):
    """Graph-owned copies of the per-call decoder inputs (encoder memory operands, frame mean)."""
    from ..engine import Act
    e = mem["enc"]
    enc = Act(e.M, e.N, *(None if t is None else torch.empty_like(t) for t in (e.f32, e.hi, e.lo)))
    return dict(enc=enc, enc_mean=torch.empty_like(mem["enc_mean"]), B=mem["B"], E=mem["E"], owner=mem["owner"])


class _DecodeGraph:
    """One captured refinement loop (all decoder passes, vocabulary statistics, refine steps and
    the candidate selection) for a fixed (batch, Smax) shape.  Inputs are copied into graph-owned
    buffers, the graph is replayed, the hypotheses are cloned out: one launch per call."""

    def __init__(self, run, opt, model, teacher_model, mem, tmem, cat, beam, S, dict_mapping, rows_hint=0):
        eng = model.engine
        self.pack_ids = (eng.pack_id, teacher_model.engine.pack_id if teacher_model is not None else -1)
        self.mem = _clone_inputs(mem)
        self.tmem = _clone_inputs(tmem) if tmem is not None else None
        self.cat = None if cat is None else torch.empty_like(cat)
        self.beam = torch.empty_like(beam)
        self.table = dict_mapping  # keeps the captured id-remap table alive
        self.load(mem, tmem, cat, beam)
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        n0 = L.launches
        with torch.cuda.graph(self.graph):
            self.hyp, self.stats = run(opt, model, teacher_model, self.mem, self.tmem, self.cat, self.beam, S, dict_mapping, rows_hint)
        self.n_launches = L.launches - n0  # kernels of ours inside the graph (replayed on every call)

    def load(self, mem, tmem, cat, beam):
        _copy_act(self.mem["enc"], mem["enc"])
        self.mem["enc_mean"].copy_(mem["enc_mean"])
        if self.tmem is not None:
            _copy_act(self.tmem["enc"], tmem["enc"])
            self.tmem["enc_mean"].copy_(tmem["enc_mean"])
        if self.cat is not None:
            self.cat.copy_(cat)
        self.beam.copy_(beam)

    def replay(self, mem, tmem, cat, beam):
        self.load(mem, tmem, cat, beam)
        self.graph.replay()
        L.launches += self.n_launches
        return self.hyp.clone()


def _run(opt, model, teacher_model, mem, tmem, cat, beam, S, dict_mapping, rows_hint=0):
    """K|V projection of the encoder memory, the refinement algorithm, the candidate selection."""
    eng = model.engine
    B, lbs = beam.shape
    mem = eng.memory(mem["enc"].f32, mem)
    if tmem is not None:
        tmem = teacher_model.engine.memory(tmem["enc"].f32, tmem)
    ref = Refiner(opt, model, teacher_model, mem, tmem, cat, beam, S, dict_mapping, rows_hint)
    tokens, lprobs = ALGORITHMS[opt.get("paradigm", "mp")](ref)
    hyp = torch.empty((B, S), dtype=torch.int64, device=beam.device)
    L.call("navc_select_best", L.ptr(tokens), L.ptr(lprobs), L.ptr(ref.lens), B, lbs, S,
           float(opt.get("beam_alpha", 1.0)), L.ptr(hyp), None, L.stream())
    return hyp, {"passes": ref.passes, "S": S, "N": ref.N, "steps": ref.n_steps, "packed": ref.packed is not None}


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
    # ONE host read: Smax, sum(len) and sum(len^2) (the last two only feed the statistics bench.py reports)
    S, rows_real, rows_sq = torch.cat([smax, beam.sum().to(torch.int32).view(1), (beam * beam).sum().to(torch.int32).view(1)]).tolist()

    mem = eng.enc_inputs(encoder_outputs["enc_output"], encoder_outputs.get("_navc"))
    tmem = None
    if teacher_model is not None and teacher_encoder_outputs is not None:
        teacher_model.engine.sync_weights()
        tmem = teacher_model.engine.enc_inputs(teacher_encoder_outputs["enc_output"], teacher_encoder_outputs.get("_navc"))
    else:
        teacher_model = None
    cat = category.contiguous() if category is not None else None
    key_map = tuple(sorted((dict_mapping or {}).items()))
    if dict_mapping:  # teacher vocabulary remap (algorithms.py:169-173) as a device lookup table
        table = torch.arange(max(dict_mapping) + 1, dtype=torch.int64)
        for k, v in dict_mapping.items():
            table[k] = v
        dict_mapping = table.to(dev)

    # mask-predict has no data-dependent host control flow: the whole loop is one CUDA
    # graph per (batch, Smax) shape.  easy-first / left-to-right read a counter per pass -> eager.
    if paradigm == "mp" and graphs_enabled(opt) and not torch.cuda.is_current_stream_capturing():
        key = (B, S, lbs, mem["E"], id(teacher_model), tuple(sorted((k, repr(v)) for k, v in opt.items()
                                                                    if k in _GRAPH_OPT_KEYS)),
               key_map, cat is None)
        pack_ids = (eng.pack_id, teacher_model.engine.pack_id if teacher_model is not None else -1)
        entry = eng.graphs.get(key)
        if entry is not None and (not isinstance(entry, _DecodeGraph) or entry.pack_ids == pack_ids):
            if entry == "warm":  # second call with this shape: capture
                live = [k for k, v in eng.graphs.items() if isinstance(v, _DecodeGraph)]
                for k in live[:max(0, len(live) - (_MAX_GRAPHS - 1))]:  # each graph owns its activations: keep a few shapes
                    del eng.graphs[k]
                entry = eng.graphs[key] = _DecodeGraph(_run, opt, model, teacher_model, mem, tmem, cat, beam, S, dict_mapping, rows_real)
            hyp = entry.replay(mem, tmem, cat, beam)
            generate.last_stats = dict(entry.stats, graph=True, rows_real=rows_real, rows_sq=rows_sq)
            return hyp, None
        eng.graphs[key] = "warm"  # first call: run eagerly (also warms up lazily initialised kernels)

    hyp, stats = _run(opt, model, teacher_model, mem, tmem, cat, beam, S, dict_mapping, rows_real)
    generate.last_stats = dict(stats, graph=False, rows_real=rows_real, rows_sq=rows_sq)
    return hyp, None


_MAX_GRAPHS = 8
_GRAPH_OPT_KEYS = ("iterations", "use_ct", "beam_alpha", "masking_decision", "no_candidate_decision", "q", "navc_packed",
                   "q_iterations", "enhance_input", "watch", "paradigm")
generate.last_stats = {}
