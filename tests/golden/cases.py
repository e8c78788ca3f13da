"""Shared definitions of the golden / parity cases (opts + seeded synthetic inputs).

Used by ``make_golden.py`` (which imports the reference to produce the fixtures), by the oracle
tests and by the GPU parity tests.  Shapes follow SURVEY.md section 8(d) / Appendix B.
"""
from __future__ import annotations

import copy

import torch

COMMON = dict(
    modality="mi", dim_a=1, dim_o=1, encoder="Encoder_HighWay", fusion="temporal_concat",
    hidden_act="gelu_new", hidden_dropout_prob=0.5, attention_probs_dropout_prob=0.0,
    layer_norm_eps=1e-5, watch=0, pos_attention=False, enhance_input=2, with_layernorm=False,
    with_category=True, num_category=20, encoder_dropout=0.5, no_encoder_bn=False, norm_type="bn",
    num_attention_heads=8, beam_alpha=1.35, paradigm="mp", use_ct=False, q=1, q_iterations=1,
    length_beam_size=3, iterations=5,
)


def make_opt(method, **kw):
    opt = copy.deepcopy(COMMON)
    opt["method"] = method
    if method == "NAB":
        opt.update(decoder="BertDecoder", decoding_type="NARFormer", visual_word_generation=False,
                   crit=["lang", "length"])
    elif method == "NACF":
        opt.update(decoder="BertDecoderDisentangled", decoding_type="NARFormer",
                   visual_word_generation=True, crit=["lang", "length"], nv_weights=[0.8, 1.0])
    elif method == "ARB":
        opt.update(decoder="BertDecoder", decoding_type="ARFormer", visual_word_generation=False,
                   crit=["lang"])
    elif method == "ARB2":
        opt.update(decoder="BertDecoderDisentangled", decoding_type="ARFormer",
                   visual_word_generation=True, crit=["lang"], nv_weights=[0.8, 1.0])
    else:
        raise ValueError(method)
    opt.update(kw)
    return opt


# BASELINE config 1 (plumbing)
def config1(**kw):
    base = dict(dim_hidden=128, num_hidden_layers_decoder=1, intermediate_size=512, dim_i=256,
                dim_m=256, n_frames=8, max_len=10, vocab_size=200)
    base.update(kw)
    return make_opt("NAB", **base)


# small NACF (2 layers) used for decode goldens with/without CT and with an ARB teacher
def small(method="NACF", **kw):
    base = dict(dim_hidden=128, num_hidden_layers_decoder=2, intermediate_size=256, dim_i=192,
                dim_m=320, n_frames=6, max_len=16, vocab_size=300, length_beam_size=4)
    base.update(kw)
    return make_opt(method, **base)


# dk = 64 variants (the head size of BASELINE configs 2-5): small with two heads, and a D=512 / 8-head "wide" model
def dk64(method="NACF", **kw):
    base = dict(num_attention_heads=2)
    base.update(kw)
    return small(method, **base)


def wide(method="NACF", **kw):
    base = dict(dim_hidden=512, num_attention_heads=8, num_hidden_layers_decoder=2, intermediate_size=1024, dim_i=256,
                dim_m=384, n_frames=20, max_len=30, vocab_size=2000, length_beam_size=5)
    base.update(kw)
    return make_opt(method, **base)


# BASELINE config 2 (headline): NACF 6-layer d512
def config2(**kw):
    base = dict(dim_hidden=512, num_hidden_layers_decoder=6, intermediate_size=2048, dim_i=2048,
                dim_m=2048, n_frames=60, max_len=30, vocab_size=10547, length_beam_size=6,
                iterations=5, use_ct=True)
    base.update(kw)
    return make_opt("NACF", **base)


def synth_inputs(opt, batch, seed=1234):
    """feats drawn modality by modality in opt['modality'] order; category randint(0,20)."""
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(batch, opt["n_frames"], opt["dim_" + ch], generator=g)
             for ch in opt["modality"].lower()]
    category = torch.randint(0, opt["num_category"], (batch, 1), generator=g)
    return feats, category


def synth_tokens(opt, batch, seed=4321, kind="nar"):
    """Training-style token tensors [B, max_len]."""
    g = torch.Generator().manual_seed(seed)
    L, V = opt["max_len"], opt["vocab_size"]
    lens = torch.randint(min(4, L - 2), L - 1, (batch,), generator=g)
    words = torch.randint(6, V, (batch, L), generator=g)
    pos = torch.arange(L).unsqueeze(0)
    inside = pos < lens.unsqueeze(1)
    if kind == "nar":
        coin = torch.rand(batch, L, generator=g) < 0.5
        tokens = torch.where(coin, torch.full_like(words, 4), words)
        tokens = torch.where(inside, tokens, torch.zeros_like(words))
        labels = torch.where(coin & inside, words, torch.zeros_like(words))
        tokens1 = torch.where(inside, torch.full_like(words, 5), torch.zeros_like(words))
        vis = torch.rand(batch, L, generator=g) < 0.4
        labels1 = torch.where(inside, torch.where(vis, words, torch.full_like(words, 4)), torch.zeros_like(words))
        length_target = torch.zeros(batch, L)
        length_target[torch.arange(batch), lens] = 1.0
        return dict(tokens=tokens, labels=labels, tokens_1=tokens1, labels_1=labels1,
                    length_target=length_target, lens=lens)
    # autoregressive: BOS w.. EOS PAD..
    seq = torch.where(inside, words, torch.zeros_like(words))
    seq[:, 0] = 2
    seq[torch.arange(batch), lens] = 3
    return dict(tokens=seq, labels=seq[:, 1:].contiguous(), lens=lens)


def seeded_state_dict(model_ctor, opt, seed=0):
    torch.manual_seed(seed)
    model = model_ctor(opt)
    model.eval()
    return model


def synth_state_dict(shapes, seed=7, scale=1.0):
    """Reference-independent seeded weights for a {name: shape} inventory (golden fixtures store the
    inventory, not the weights).  Matrices ~N(0, 0.08), biases ~N(0, 0.05), norm weights ~1+N(0,0.1),
    BN running_var in [0.5, 1.5]; embedding row PAD is zero as nn.Embedding(padding_idx=0) keeps it.
    ``scale`` multiplies the matrix standard deviation (wide models: keeps the softmax away from saturation)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(3, dtype=torch.long)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif name.endswith("running_mean"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif ("LayerNorm" in name or ".bn" in name or ".ln" in name) and name.endswith("weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif "embeddings" in name:
            w = 0.5 * torch.randn(shape, generator=g)
            if "word_embeddings" in name:
                w[0].zero_()
            sd[name] = w
        else:
            sd[name] = (0.08 * scale) * torch.randn(shape, generator=g)
    return sd
