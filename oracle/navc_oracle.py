"""CPU oracle for the NA video-captioning hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch *restatement* (plain torch fp32 on the CPU, functional, driven by a
``state_dict``) of the algorithm implemented by the reference repository
yangbang18/Non-Autoregressive-Video-Captioning.  It exists to check the CUDA product path; it is
never imported by the product package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity pinning: the reference ships no tests / golden vectors for this path, so
the oracle is pinned against (a) the unmodified reference imported from ``/root/reference`` in
the build container (``tests/test_oracle_vs_reference.py``; skipped where the reference is not
mounted) and (b) golden vectors produced from the reference by ``tests/golden/make_golden.py``
and committed under ``tests/golden/`` (``tests/test_oracle_golden.py``).

Each function cites the reference file:line it restates.

Tie-breaking: the reference calls ``topk(sorted=False)`` whose tie order is
implementation-defined.  The oracle *defines* lowest-index-first; comparisons against the
reference assert that no tie occurs at the k-th boundary.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

# config/Constants.py:1-6
PAD, UNK, BOS, EOS, MASK, VIS = 0, 1, 2, 3, 4, 5
MASK_FILL = -10e6  # models/bert.py:161  (== -1e7, applied after the 1/sqrt(dk) scale)


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def _linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _layernorm(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def gelu_new(x):
    """models/bert.py:12-13 (tanh form)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def activation(name):
    """models/bert.py:9-19 ACT2FN."""
    if name == "gelu_new":
        return gelu_new
    if name == "gelu":
        return lambda x: x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if name == "relu":
        return F.relu
    if name == "swish":
        return lambda x: x * torch.sigmoid(x)
    raise ValueError(name)


def is_nar(opt) -> bool:
    return opt["decoding_type"] == "NARFormer"


def decoder_prefix(opt) -> str:
    """BertDecoderDisentangled nests the shared decoder under ``.bert`` (models/Decoder.py:181-186)."""
    return "decoder.bert" if opt["decoder"] == "BertDecoderDisentangled" else "decoder"


# ----------------------------------------------------------------------------------------------
# encoder side
# ----------------------------------------------------------------------------------------------
def encoder_stream(sd, prefix, feats, drop_p=0.0, training=False):
    """One modality stream: Linear -> HighWay -> Dropout (models/Encoder.py:19-25, 62-66)."""
    x = _linear(sd, prefix + ".0", feats)
    y = torch.tanh(_linear(sd, prefix + ".1.w1", x))
    if (prefix + ".1.w2.weight") in sd:
        g = torch.sigmoid(_linear(sd, prefix + ".1.w2", x))
        out = g * x + (1.0 - g) * y
    else:
        out = x + y
    return F.dropout(out, drop_p, training)


def encode(sd, opt, feats: Sequence[torch.Tensor], training=False, bn_state=None):
    """models/seq2seq.py:35-63 -> Encoder.py:47-59 -> joint_representation.py:24-53 -> Predictor.py:23-30.

    ``training`` selects batch statistics for BatchNorm (dropout is applied with the opt
    probabilities, so gradient-parity tests set those to 0).  ``bn_state`` (optional dict) receives
    the updated running statistics, as nn.BatchNorm1d would store them.
    """
    modality = opt["modality"].lower()
    assert len(modality) == len(feats)
    outs, hiddens = [], []
    for ch, f in zip(modality, feats):
        o = encoder_stream(sd, "encoder.Encoder_%s" % ch.upper(), f,
                           opt.get("encoder_dropout", 0.5), training)
        outs.append(o)
        hiddens.append(o.mean(1))
    enc_hidden = torch.stack(hiddens, 0).mean(0)

    fusion = opt.get("fusion", "temporal_concat")
    if fusion == "none":
        enc_output = torch.cat(outs, 1)
    else:
        if fusion == "addition":
            # joint_representation.py:37-41: after the stack/mean the reference asserts
            # len(tensor) == len(norm_list) and fails for any batch != 1 -> not a usable path.
            raise NotImplementedError("fusion='addition' is broken in the reference (joint_representation.py:41)")
        if not opt["no_encoder_bn"]:
            is_bn = opt.get("norm_type", "bn").lower() == "bn"
            for i in range(len(outs)):
                name = "joint_representation_learner.%s%d" % ("bn" if is_bn else "ln", i)
                if is_bn:
                    b, t, d = outs[i].shape
                    rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
                    if training and bn_state is not None:
                        rm, rv = rm.clone(), rv.clone()
                    flat = F.batch_norm(outs[i].contiguous().view(b * t, d), rm, rv,
                                        sd[name + ".weight"], sd[name + ".bias"],
                                        training, 0.1, 1e-5)
                    if training and bn_state is not None:
                        bn_state[name + ".running_mean"] = rm
                        bn_state[name + ".running_var"] = rv
                    outs[i] = flat.view(b, t, d)
                else:
                    outs[i] = _layernorm(sd, name, outs[i], 1e-5)
        enc_output = torch.cat(outs, 1) if fusion == "temporal_concat" else outs[0]

    results = {}
    if "auxiliary_task_predictor.layers.0.net.0.weight" in sd:
        h = _linear(sd, "auxiliary_task_predictor.layers.0.net.0", enc_output.mean(1))
        h = F.dropout(F.relu(h), opt["hidden_dropout_prob"], training)
        h = _linear(sd, "auxiliary_task_predictor.layers.0.net.3", h)
        results["pred_length"] = torch.log_softmax(h, dim=-1)
    results["enc_output"] = enc_output
    results["enc_hidden"] = enc_hidden
    return results


# ----------------------------------------------------------------------------------------------
# decoder side
# ----------------------------------------------------------------------------------------------
def self_attention_mask(tgt_seq, decoding_type, watch=0):
    """models/Decoder.py:13-39, 105-124.  True = masked.  [N,S,S] bool."""
    n, s = tgt_seq.shape
    keypad = tgt_seq.eq(PAD).unsqueeze(1).expand(-1, s, -1)
    if decoding_type == "NARFormer":
        return keypad
    if decoding_type == "SelfMask":
        eye = torch.eye(s, dtype=torch.bool, device=tgt_seq.device)
        return keypad | eye.unsqueeze(0)
    future = torch.triu(torch.ones(s, s, dtype=torch.bool, device=tgt_seq.device), diagonal=1)
    if watch != 0 and s >= watch:
        future = future | torch.tril(torch.ones(s, s, dtype=torch.bool, device=tgt_seq.device),
                                     diagonal=-watch)
    return keypad | future.unsqueeze(0)


def multi_head_attention(sd, prefix, q_in, kv_in, mask, n_head):
    """models/bert.py:139-179.  Returns (context [N,Sq,D], probs [H,N,Sq,Sk])."""
    n, sq, d = q_in.shape
    sk = kv_in.shape[1]
    dk = d // n_head
    q = _linear(sd, prefix + ".query", q_in).view(n, sq, n_head, dk).permute(2, 0, 1, 3)
    k = _linear(sd, prefix + ".key", kv_in).view(n, sk, n_head, dk).permute(2, 0, 1, 3)
    v = _linear(sd, prefix + ".value", kv_in).view(n, sk, n_head, dk).permute(2, 0, 1, 3)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dk)  # [H,N,Sq,Sk]
    if mask is not None:
        scores = scores.masked_fill(mask.unsqueeze(0), MASK_FILL)
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v)  # [H,N,Sq,dk]
    ctx = ctx.permute(1, 2, 0, 3).contiguous().view(n, sq, d)
    return ctx, probs


def attention_block(sd, prefix, q_in, kv_in, mask, n_head, eps, drop_p, training):
    """BertAttention = MHA + BertSelfOutput (models/bert.py:192-215)."""
    ctx, probs = multi_head_attention(sd, prefix + ".self", q_in, kv_in, mask, n_head)
    h = F.dropout(_linear(sd, prefix + ".output.dense", ctx), drop_p, training) + q_in
    if (prefix + ".output.LayerNorm.weight") in sd:
        h = _layernorm(sd, prefix + ".output.LayerNorm", h, eps)
    return h, probs


def bert_layer(sd, opt, prefix, x, non_pad, slf_mask, enc_output, training=False):
    """models/bert.py:262-303 (decoder layer; pos_attention unsupported as in methods.yaml)."""
    H, eps, p = opt["num_attention_heads"], opt["layer_norm_eps"], opt["hidden_dropout_prob"]
    a, p_self = attention_block(sd, prefix + ".attention", x, x, slf_mask, H, eps, p, training)
    a = a * non_pad
    # cross attention mask is built from an all-ones dummy source => nothing masked (Decoder.py:127-128)
    c, p_cross = attention_block(sd, prefix + ".attend_to_enc_output", a, enc_output, None, H, eps, p, training)
    c = c * non_pad
    h = activation(opt["hidden_act"])(_linear(sd, prefix + ".intermediate.dense", c))
    y = F.dropout(_linear(sd, prefix + ".output.dense", h), p, training) + c  # bert.py:240-243
    if (prefix + ".output.LayerNorm.weight") in sd:
        y = _layernorm(sd, prefix + ".output.LayerNorm", y, eps)
    y = F.dropout(y, p, training)  # second dropout, bert.py:247
    y = y * non_pad
    return y, (p_self, p_cross)


def embeddings(sd, opt, prefix, tgt_seq, category, additional, training=False):
    """models/bert.py:70-96."""
    n, s = tgt_seq.shape
    e = F.embedding(tgt_seq, sd[prefix + ".word_embeddings.weight"])
    e = e + sd[prefix + ".position_embeddings.weight"][:s].unsqueeze(0)
    if (prefix + ".category_embeddings.weight") in sd:
        e = e + F.embedding(category, sd[prefix + ".category_embeddings.weight"]).expand(-1, s, -1)
    if additional is not None:
        e = e + additional
    e = _layernorm(sd, prefix + ".LayerNorm", e, opt["layer_norm_eps"])
    return F.dropout(e, opt["hidden_dropout_prob"], training)


def decoder_forward(sd, opt, tgt_seq, enc_output, category, decoding_type=None,
                    output_attentions=False, training=False, prefix=None):
    """BertDecoder.forward (models/Decoder.py:96-178).  Returns (hidden [N,S,D], embs [N,D], attns)."""
    prefix = prefix or decoder_prefix(opt)
    decoding_type = decoding_type or opt["decoding_type"]
    assert not opt.get("pos_attention", False)
    slf_mask = self_attention_mask(tgt_seq, decoding_type, opt.get("watch", 0))
    non_pad = tgt_seq.ne(PAD).float().unsqueeze(-1)
    additional = None
    if decoding_type == "NARFormer":
        ei = opt.get("enhance_input", 2)
        if ei == 2:
            additional = enc_output.mean(1, keepdim=True).expand(-1, tgt_seq.shape[1], -1)
        elif ei != 0:
            raise ValueError("enhance_input=1 is broken in the reference (SURVEY.md 8c); unsupported")
    x = embeddings(sd, opt, prefix + ".embedding", tgt_seq, category, additional, training)
    attns = []
    for l in range(opt["num_hidden_layers_decoder"]):
        x, a = bert_layer(sd, opt, "%s.layer.%d" % (prefix, l), x, non_pad, slf_mask, enc_output, training)
        attns.append(a)
    embs = x.sum(1) / non_pad.sum(1)
    return x, embs, (attns if output_attentions else None)


def vocab_logits(sd, hidden):
    """Seq2Seq.tgt_word_prj (models/__init__.py:83; bias only with tie_weights, seq2seq.py:30-33)."""
    return _linear(sd, "tgt_word_prj", hidden)


def model_forward(sd, opt, feats, tgt_tokens, category, training=False, bn_state=None):
    """Seq2Seq.forward_NARFormer / forward_ARFormer (models/seq2seq.py:86-140)."""
    results = encode(sd, opt, feats, training, bn_state)
    seqs = tgt_tokens if isinstance(tgt_tokens, (list, tuple)) else [tgt_tokens]
    if not is_nar(opt):
        seqs = [t[:, :-1] for t in seqs]
    logprobs = []
    for t in seqs:
        h, _, _ = decoder_forward(sd, opt, t, results["enc_output"], category, training=training)
        logprobs.append(torch.log_softmax(vocab_logits(sd, h), dim=-1))
    results["tgt_word_logprobs"] = logprobs
    return results


def criterion(opt, results, labels, length_target=None):
    """misc/crit.py:40-46, 62-84, 156-181, 223: masked NLL sum / batch (+ KLDiv 'mean' on length)."""
    labs = labels if isinstance(labels, (list, tuple)) else [labels] * len(results["tgt_word_logprobs"])
    weights = opt.get("nv_weights", [0.8, 1.0]) if opt.get("visual_word_generation", False) else [1.0] * len(labs)
    bsz = results["tgt_word_logprobs"][0].shape[0]
    loss = 0.0
    for w, lp, lab in zip(weights, results["tgt_word_logprobs"], labs):
        nll = F.nll_loss(lp.reshape(-1, lp.shape[-1]), lab.reshape(-1), reduction="none")
        loss = loss + w * (nll * lab.reshape(-1).ne(PAD).float()).sum() / bsz
    if length_target is not None and "pred_length" in results:
        # nn.KLDivLoss() default reduction 'mean' = mean over all elements
        kl = F.kl_div(results["pred_length"], length_target, reduction="mean")
        loss = loss + kl
    return loss


# ----------------------------------------------------------------------------------------------
# iterative-refinement decoding
# ----------------------------------------------------------------------------------------------
def enlarge(x, k):
    """misc/utils.py:205-213: candidate-major repeat (row = b*k + j)."""
    return x.unsqueeze(1).expand(x.shape[0], k, *x.shape[1:]).reshape(x.shape[0] * k, *x.shape[1:])


def length_beam(pred_length, lbs, length_bias, max_len):
    """decoding/na_generate.py:116-135."""
    beam = pred_length.topk(lbs, dim=1)[1] + length_bias
    return beam.clamp(min=4, max=max_len - 1)


def k_smallest_mask(probs, k):
    """select_worst (decoding/algorithms.py:206-215) with lowest-index-first ties; k>=1 enforced."""
    k = k.clamp(min=1)
    order = torch.sort(probs, dim=1, stable=True)[1]
    rank = torch.empty_like(order)
    rank.scatter_(1, order, torch.arange(probs.shape[1]).expand_as(order))
    return rank < k.unsqueeze(1)


def k_largest_mask(probs, k):
    """select_most_confidence (algorithms.py:297-309) with lowest-index-first ties; k may be 0."""
    order = torch.sort(probs, dim=1, descending=True, stable=True)[1]
    rank = torch.empty_like(order)
    rank.scatter_(1, order, torch.arange(probs.shape[1]).expand_as(order))
    return rank < k.unsqueeze(1)


def boundary_gap_smallest(probs, k):
    """gap between the k-th and (k+1)-th smallest value per row (inf if k == S).  For tie asserts."""
    k = k.clamp(min=1)
    srt = torch.sort(probs, dim=1)[0]
    s = probs.shape[1]
    kth = srt.gather(1, (k - 1).clamp(max=s - 1).unsqueeze(1)).squeeze(1)
    nxt = srt.gather(1, k.clamp(max=s - 1).unsqueeze(1)).squeeze(1)
    gap = (nxt - kth).abs()
    gap[k >= s] = float("inf")
    return gap


class Decoder:
    """Binds (sd, opt) for the decode loop; ``stats`` records margins for the parity tests."""

    def __init__(self, sd, opt, teacher=None):
        self.sd, self.opt = sd, opt
        self.teacher = teacher  # (sd, opt) or None
        self.passes = 0
        self.min_top2_gap = float("inf")
        self.min_select_gap = float("inf")
        # per candidate row, in log units (logit / log-prob differences): the smallest margin of any decision
        # whose outcome reaches the row's final tokens -- the argmax of a position that is merged, and the
        # k-th / (k+1)-th boundary of every selection.  Parity reports use it to tell a real mismatch from a
        # decision that sits inside the arithmetic error of the mode under test.
        self.row_margin = None

    def _row_min(self, values, valid):
        """row_margin[n] = min(row_margin[n], min over valid positions of values[n, :])."""
        v = values.masked_fill(~valid, float("inf")).min(dim=1)[0]
        self.row_margin = v if self.row_margin is None else torch.minimum(self.row_margin, v)

    # algorithms.py:7-15, 143-167
    def na_pass(self, tokens, enc_output, category, pad_mask, used=None):
        h, _, _ = decoder_forward(self.sd, self.opt, tokens, enc_output, category, output_attentions=False)
        logits = vocab_logits(self.sd, h)
        probs = torch.softmax(logits, dim=-1)
        p, idx = probs.max(dim=-1)
        top2 = logits.topk(2, dim=-1)[0]
        # positions whose *input* token is PAD have an all-zero hidden state => all logits equal;
        # first-index argmax (= PAD) is the defined result there, so they are not counted as ties.
        live = ~(pad_mask | tokens.eq(PAD))
        gap = (top2[..., 0] - top2[..., 1])[live]
        if gap.numel():
            self.min_top2_gap = min(self.min_top2_gap, gap.min().item())
        self._row_min(top2[..., 0] - top2[..., 1], live if used is None else (live & used))
        idx = idx.masked_fill(pad_mask, PAD)
        p = p.masked_fill(pad_mask, 1.0)
        self.passes += 1
        return idx, p

    # algorithms.py:136-141
    def ct_pass(self, tokens, enc_output, category, pad_mask):
        canvas = tokens.masked_fill(tokens.eq(MASK), VIS)
        idx, p = self.na_pass(canvas, enc_output, category, pad_mask)
        p = p.masked_fill(idx.eq(MASK), 0.0)
        return idx, p

    # algorithms.py:175-204
    def teacher_probs(self, tokens, t_enc_output, category, pad_mask, is_last, dict_mapping=None):
        ones = torch.ones(tokens.shape, dtype=torch.float32)
        if self.teacher is None:
            return ones
        if is_last and self.opt.get("no_candidate_decision", False):
            return ones
        if (not is_last) and not self.opt.get("masking_decision", False):
            return ones
        tsd, topt = self.teacher
        toks = tokens
        if dict_mapping:
            toks = tokens.clone().apply_(lambda t: dict_mapping[int(t)])
        shifted = torch.cat([torch.full((toks.shape[0], 1), BOS, dtype=toks.dtype), toks[:, :-1]], 1)
        h, _, _ = decoder_forward(tsd, topt, shifted, t_enc_output, category)
        probs = torch.softmax(vocab_logits(tsd, h), dim=-1)
        p = probs.gather(2, toks.unsqueeze(2)).squeeze(2)
        return p.masked_fill(pad_mask, 1.0)

    def _select_worst(self, probs, k):
        self.min_select_gap = min(self.min_select_gap, boundary_gap_smallest(probs, k).min().item())
        lg = boundary_gap_smallest(probs.clamp_min(1e-38).log(), k)
        self.row_margin = lg if self.row_margin is None else torch.minimum(self.row_margin, lg)
        return k_smallest_mask(probs, k)

    # algorithms.py:231-273
    def mask_predict(self, tokens, enc_output, t_enc_output, category):
        opt = self.opt
        pad_mask = tokens.eq(PAD)
        lens = tokens.shape[1] - pad_mask.sum(1)
        use_ct = opt.get("use_ct", False)
        T = opt.get("iterations", 5) + (1 if use_ct else 0)
        if use_ct:
            tok, prob = self.ct_pass(tokens, enc_output, category, pad_mask)
        else:
            tok, prob = self.na_pass(tokens, enc_output, category, pad_mask)
        for t in range(1, T):
            pt = self.teacher_probs(tok, t_enc_output, category, pad_mask, is_last=False)
            if use_ct and t == 1:
                mask = tok.eq(MASK)
            else:
                ratio = 1.0 - (t / T)
                k = (lens.float() * ratio).long()
                mask = self._select_worst(prob * pt, k)
            tok = tok.masked_fill(mask, MASK)
            ntok, nprob = self.na_pass(tok, enc_output, category, pad_mask, used=mask)
            tok = torch.where(mask, ntok, tok)
            prob = torch.where(mask, nprob, prob)
        pt = self.teacher_probs(tok, t_enc_output, category, pad_mask, is_last=True)
        return tok, (prob * pt).log()

    def _refine_tail(self, tok, prob, lens, visual_mask, enc_output, category, pad_mask):
        # shared refinement tail of EasyFirst / Left2Right (algorithms.py:326-339, 398-411)
        Tq = self.opt.get("q_iterations", 1)
        for i in range(Tq):
            if i == 0 and self.opt.get("use_ct", False):
                mask = visual_mask
            else:
                ratio = 0.4 * (1.0 - (i / Tq))
                k = (lens.float() * ratio).long()
                mask = self._select_worst(prob, k)
            tok = tok.masked_fill(mask, MASK)
            ntok, nprob = self.na_pass(tok, enc_output, category, pad_mask, used=mask)
            tok = torch.where(mask, ntok, tok)
            prob = torch.where(mask, nprob, prob)
        return tok, prob

    def _start(self, tokens, enc_output, category, pad_mask):
        if self.opt.get("use_ct", False):
            tok, prob = self.ct_pass(tokens, enc_output, category, pad_mask)
            visual = tok.ne(MASK) & tok.ne(PAD)
        else:
            tok = tokens.clone()
            prob = pad_mask.float()  # 0 everywhere, 1 at pads
            visual = None
        return tok, prob, visual

    # algorithms.py:354-418
    def easy_first(self, tokens, enc_output, t_enc_output, category):
        pad_mask = tokens.eq(PAD)
        lens = tokens.shape[1] - pad_mask.sum(1)
        q = self.opt.get("q", 1)
        tok, prob, visual = self._start(tokens, enc_output, category, pad_mask)
        prev = 0
        while True:
            mask = tok.eq(MASK)
            remain = int(mask.sum())
            if remain == 0 or remain == prev:
                break
            prev = remain
            ntok, nprob = self.na_pass(tok, enc_output, category, pad_mask, used=mask)
            cand = nprob.masked_fill(~mask, 0.0)
            k = mask.sum(1).clamp(max=q)
            # boundary-gap bookkeeping for the k largest
            srt = torch.sort(cand, dim=1, descending=True)[0]
            s = cand.shape[1]
            kk = k.clamp(min=1)
            hi_, lo_ = srt.gather(1, (kk - 1).unsqueeze(1)).squeeze(1), srt.gather(1, kk.clamp(max=s - 1).unsqueeze(1)).squeeze(1)
            gap = (hi_ - lo_).abs()
            # per row, in log units; rows that commit every remaining position (k == remaining) decide nothing
            lg = (hi_.clamp_min(1e-38).log() - lo_.clamp_min(1e-38).log()).abs()
            lg[(k == 0) | (k >= mask.sum(1)) | (k >= s)] = float("inf")
            self.row_margin = lg if self.row_margin is None else torch.minimum(self.row_margin, lg)
            gap = gap[(k > 0) & (k < s)]
            if gap.numel():
                self.min_select_gap = min(self.min_select_gap, gap.min().item())
            commit = k_largest_mask(cand, k)
            tok = torch.where(commit, ntok, tok)
            prob = torch.where(commit, cand, prob)
        tok, prob = self._refine_tail(tok, prob, lens, visual, enc_output, category, pad_mask)
        pt = self.teacher_probs(tok, t_enc_output, category, pad_mask, is_last=True)
        return tok, (prob * pt).log()

    # algorithms.py:282-344
    def left_to_right(self, tokens, enc_output, t_enc_output, category):
        pad_mask = tokens.eq(PAD)
        n, s = tokens.shape
        lens = s - pad_mask.sum(1)
        q = self.opt.get("q", 1)
        tok, prob, visual = self._start(tokens, enc_output, category, pad_mask)
        # ordinal (0-based) of each originally-masked position among the masked positions of its row
        in_len = torch.arange(s).unsqueeze(0) < lens.unsqueeze(1)
        masked0 = tok.eq(MASK) & in_len
        ordinal = masked0.long().cumsum(1) - 1
        for cur in range(0, s, q):
            mask = masked0 & (ordinal >= cur) & (ordinal < cur + q)
            if int(mask.sum()) == 0:
                break
            tok = tok.masked_fill(mask, MASK)
            ntok, nprob = self.na_pass(tok, enc_output, category, pad_mask, used=mask)
            tok = torch.where(mask, ntok, tok)
            prob = torch.where(mask, nprob, prob)
        tok, prob = self._refine_tail(tok, prob, lens, visual, enc_output, category, pad_mask)
        pt = self.teacher_probs(tok, t_enc_output, category, pad_mask, is_last=True)
        return tok, (prob * pt).log()


def generate(sd, opt, encoder_outputs, category, teacher=None, teacher_encoder_outputs=None,
             length_bias=0, return_details=False):
    """decoding/na_generate.py:14-108 (default path: no gold lengths, no collection)."""
    pred_length = encoder_outputs["pred_length"]
    bsz = pred_length.shape[0]
    lbs = opt["length_beam_size"]
    beam = length_beam(pred_length, lbs, length_bias, opt["max_len"])
    smax = int(beam.max())
    pos = torch.arange(smax).view(1, 1, smax)
    tokens = torch.where(pos < beam.unsqueeze(-1), torch.full((1,), MASK), torch.full((1,), PAD))
    tokens = tokens.view(bsz * lbs, smax)
    enc_output = enlarge(encoder_outputs["enc_output"], lbs)
    cat = enlarge(category, lbs)
    t_enc = enlarge(teacher_encoder_outputs["enc_output"], lbs) if teacher_encoder_outputs is not None else None
    dec = Decoder(sd, opt, teacher if teacher_encoder_outputs is not None else None)
    algo = {"mp": dec.mask_predict, "ef": dec.easy_first, "l2r": dec.left_to_right}[opt.get("paradigm", "mp")]
    tok, lprobs = algo(tokens, enc_output, t_enc, cat)
    tok = tok.view(bsz, lbs, smax)
    lprobs = lprobs.view(bsz, lbs, smax)
    score = lprobs.sum(-1) / (beam.float() ** opt.get("beam_alpha", 1.0))
    best = score.max(-1)[1]
    hyp = tok.gather(1, best.view(bsz, 1, 1).expand(bsz, 1, smax)).squeeze(1)
    if return_details:
        # margin between the winner and the best candidate with a *different* hypothesis (duplicate
        # lengths after clamping give identical candidates, which are not ties that matter)
        same = (tok == hyp.unsqueeze(1)).all(-1)
        other = score.masked_fill(same, float("-inf")).max(-1)[0]
        cand_gaps = score.max(-1)[0] - other
        cand_gap = cand_gaps.min().item()
        # per video, in log units: every decision of every candidate row, the candidate choice, and the boundary
        # of the length beam (k-th vs (k+1)-th most likely length; a different beam changes the whole video)
        srt = pred_length.sort(dim=1, descending=True)[0]
        beam_gap = (srt[:, lbs - 1] - srt[:, lbs]) if srt.shape[1] > lbs else torch.full((bsz,), float("inf"))
        video_margin = torch.minimum(torch.minimum(dec.row_margin.view(bsz, lbs).min(1)[0], cand_gaps), beam_gap)
        return hyp, {"beam": beam, "tokens": tok, "lprobs": lprobs, "score": score, "best": best,
                     "passes": dec.passes, "min_top2_gap": dec.min_top2_gap,
                     "min_select_gap": dec.min_select_gap, "min_candidate_gap": cand_gap,
                     "video_margin": video_margin, "beam_gap": beam_gap}
    return hyp


def translate(sd, opt, feats, category, teacher=None, length_bias=0, return_details=False):
    """model.encode + Translator.translate_batch under no_grad (misc/run.py:130-141; Translator.py:163-185)."""
    with torch.no_grad():
        enc = encode(sd, opt, feats)
        t_enc = encode(teacher[0], teacher[1], feats) if teacher is not None else None
        return generate(sd, opt, enc, category, teacher, t_enc, length_bias, return_details)


# ----------------------------------------------------------------------------------------------
# autoregressive beam search (models/Translator.py:94-161 + models/Beam.py), per video
# ----------------------------------------------------------------------------------------------
def ar_beam_search(sd, opt, encoder_outputs, category):
    """Restates Translator.translate_batch_ARFormer with Beam (ARFormer branch): returns
    (hyps[b][n] token lists without BOS, scores[b][n]) for n < topk."""
    n_bm, max_len = int(opt["beam_size"]), int(opt["max_len"])
    n_sents = max(n_bm, int(opt.get("topk", 1)))
    enc = encoder_outputs["enc_output"]
    B = enc.shape[0]
    all_h, all_s = [], []
    for b in range(B):
        e = enc[b:b + 1].expand(n_bm, -1, -1)
        c = category[b:b + 1].expand(n_bm, -1)
        scores = torch.zeros(n_bm)
        next_ys = [torch.full((n_bm,), PAD, dtype=torch.long)]
        next_ys[0][0] = BOS
        prev_ks, finished, done = [], [], False
        for t in range(1, max_len):
            if len(next_ys) == 1:
                seq = next_ys[0].unsqueeze(1)
            else:
                hyps = []
                for k in range(n_bm):
                    h, kk = [], k
                    for j in range(len(prev_ks) - 1, -1, -1):
                        h.append(int(next_ys[j + 1][kk]))
                        kk = int(prev_ks[j][kk])
                    hyps.append([BOS] + h[::-1])
                seq = torch.tensor(hyps, dtype=torch.long)
            hid, _, _ = decoder_forward(sd, opt, seq, e, c)
            logp = torch.log_softmax(vocab_logits(sd, hid[:, -1, :]), dim=1)  # Translator.py:108-114
            V = logp.shape[1]
            if prev_ks:
                lk = logp + scores.unsqueeze(1)
                lk[next_ys[-1].eq(EOS)] = -1e20
            else:
                lk = logp[0]
            best, ids = lk.reshape(-1).topk(n_bm, 0, True, True)
            scores = best
            pk = ids // V
            prev_ks.append(pk)
            next_ys.append(ids - pk * V)
            for i in range(n_bm):
                if int(next_ys[-1][i]) == EOS:
                    finished.append([float(scores[i]), len(next_ys) - 1, i])
                    if len(finished) >= n_sents:
                        done = True
                        break
            if not done and len(next_ys) == max_len:
                done = True
                if not finished:
                    for i in range(n_bm):
                        finished.append([float(scores[i]), len(next_ys) - 1, i])
            if done:
                break
        alpha = opt.get("beam_alpha", 1.0)
        items = sorted(([sc / (t ** alpha), t, k] for sc, t, k in finished), key=lambda a: -a[0])[:int(opt.get("topk", 1))]
        hs = []
        for _, t, k in items:
            h = []
            for j in range(t - 1, -1, -1):
                h.append(int(next_ys[j + 1][k]))
                k = int(prev_ks[j][k])
            hs.append(h[::-1])
        all_h.append(hs)
        all_s.append([it[0] for it in items])
    return all_h, all_s
