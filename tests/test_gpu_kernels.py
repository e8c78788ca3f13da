"""Kernel-level parity through the C ABI (ctypes) against the oracle / plain torch fp32 on the same
seeded inputs.  Tolerances: fp32 CUDA-core kernels 1e-5 relative (accumulation order only);
bf16x3 tensor-core GEMM 2e-5 of the output scale; bf16 GEMM 1e-2 (operand rounding);
integer / index outputs bit-exact."""
import math

import pytest
import torch

import navc_b200
from navc_b200 import _lib as L
from oracle import navc_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.fixture(scope="module", autouse=True)
def _init():
    L.ensure_init(DEV)
    torch.cuda.set_device(0)


def g(seed):
    return torch.Generator().manual_seed(seed)


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def run_linear(x, w, b=None, act=0, residual=None, row_tokens=None, mode="f32", want_bf=False):
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=DEV)
    ohi = torch.empty((M, N), dtype=torch.bfloat16, device=DEV) if want_bf else None
    olo = torch.empty((M, N), dtype=torch.bfloat16, device=DEV) if want_bf else None
    ep = L.Epilogue(L.ptr(b), L.ptr(residual), L.ptr(row_tokens), act, N if residual is not None else 0,
                    L.ptr(out), L.ptr(ohi), L.ptr(olo), N, 0)
    if mode == "f32":
        L.call("navc_linear_f32", L.ptr(x), K, L.ptr(w), K, M, N, K, ep, L.stream())
    else:
        xh, xl = split(x)
        wh, wl = split(w)
        L.call("navc_linear_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(xh), L.ptr(xl), K,
               L.ptr(wh), L.ptr(wl), K, M, N, K, ep, L.stream())
    torch.cuda.synchronize()
    return out, ohi, olo


SHAPES = [(5, 30, 64), (128, 256, 64), (257, 384, 128), (1000, 512, 512), (300, 200, 128), (129, 1000, 192),
          (2048, 2048, 512), (4096, 512, 2048)]


@pytest.mark.parametrize("mode,tol", [("f32", 2e-6), ("bf16x3", 2e-5), ("bf16", 1e-2)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_plain(mode, tol, M, N, K):
    x = torch.randn(M, K, generator=g(1)).to(DEV)
    w = (torch.randn(N, K, generator=g(2)) / math.sqrt(K)).to(DEV)
    ref = (x.double() @ w.double().t()).float()
    out, _, _ = run_linear(x, w, mode=mode)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, err


@pytest.mark.parametrize("mode,tol", [("f32", 1e-5), ("bf16x3", 3e-5)])
@pytest.mark.parametrize("act", ["none", "gelu_new", "gelu", "relu", "swish"])
def test_linear_epilogue(mode, tol, act):
    M, N, K = 333, 512, 256
    x = torch.randn(M, K, generator=g(3)).to(DEV)
    w = (torch.randn(N, K, generator=g(4)) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g(5)).to(DEV)
    res = torch.randn(M, N, generator=g(6)).to(DEV)
    toks = torch.randint(0, 3, (M,), generator=g(7)).to(DEV)
    out, ohi, olo = run_linear(x, w, b, L.ACT[act], res, toks, mode=mode, want_bf=True)
    y = (x.double() @ w.double().t()).float() + b
    y = O.activation(act)(y) if act != "none" else y
    y = (y + res) * toks.ne(0).float().unsqueeze(1)
    assert (out - y).abs().max().item() < tol * max(1.0, y.abs().max().item())
    assert (out[toks == 0] == 0).all()
    # hi/lo split of the stored value is exact by construction
    hi, lo = split(out)
    assert torch.equal(ohi, hi) and torch.equal(olo, lo)
    assert ((ohi.float() + olo.float()) - out).abs().max().item() <= 2 ** -16 * out.abs().max().item()


@pytest.mark.parametrize("mode", ["f32", "bf16x3", "bf16"])
@pytest.mark.parametrize("M,V,K,bias", [(77, 200, 128, False), (500, 10547, 512, False), (260, 1031, 128, True)])
def test_vocab_partials(mode, M, V, K, bias):
    h = torch.randn(M, K, generator=g(8)).to(DEV)
    w = (torch.randn(V, K, generator=g(9)) / math.sqrt(K)).to(DEV)
    b = torch.randn(V, generator=g(10)).to(DEV) if bias else None
    tgt = torch.randint(0, V, (M,), generator=g(11)).to(DEV)
    tc = mode != "f32"
    tile = L._lib.navc_vocab_tile(1 if tc else 0)
    nt = (V + tile - 1) // tile
    pm = torch.empty((M, nt), device=DEV)
    ps = torch.empty((M, nt), device=DEV)
    pi = torch.empty((M, nt), dtype=torch.int32, device=DEV)
    tl = torch.empty((M,), device=DEV)
    if tc:
        hh, hl = split(h)
        wh, wl = split(w)
        L.call("navc_vocab_partials_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(hh), L.ptr(hl), K,
               L.ptr(wh), L.ptr(wl), K, L.ptr(b), M, V, K, L.ptr(pm), L.ptr(ps), L.ptr(pi), L.ptr(tgt), L.ptr(tl), L.stream())
    else:
        L.call("navc_vocab_partials_f32", L.ptr(h), K, L.ptr(w), K, L.ptr(b), M, V, K, L.ptr(pm), L.ptr(ps), L.ptr(pi),
               L.ptr(tgt), L.ptr(tl), L.stream())
    torch.cuda.synchronize()
    logits = (h.double() @ w.double().t()).float()
    if bias:
        logits = logits + b
    gm = pm.max(1)[0]
    total = (ps * torch.exp(pm - gm.unsqueeze(1))).sum(1)
    prob = 1.0 / total
    ref_p, ref_i = torch.softmax(logits, -1).max(-1)
    tol = {"f32": 1e-5, "bf16x3": 5e-5, "bf16": 5e-2}[mode]
    assert ((prob - ref_p).abs() / ref_p).max().item() < tol
    tile_of_max = pm.argmax(1)
    idx = pi.gather(1, tile_of_max.unsqueeze(1)).squeeze(1).long()
    top2 = logits.topk(2, -1)[0]
    clear = (top2[:, 0] - top2[:, 1]) > (1e-4 if mode != "bf16" else 5e-2)
    assert torch.equal(idx[clear], ref_i[clear])
    ref_tl = logits.gather(1, tgt.unsqueeze(1)).squeeze(1)
    assert (tl - ref_tl).abs().max().item() < tol * 10 * max(1.0, ref_tl.abs().max().item())


def test_vocab_ties_resolve_to_lowest_index():
    # all-zero hidden rows => all logits equal => argmax must be column 0 (PAD), prob = 1/V
    M, V, K = 40, 700, 64
    h = torch.zeros(M, K, device=DEV)
    w = torch.randn(V, K, generator=g(12)).to(DEV)
    for tc in (0, 1):
        tile = L._lib.navc_vocab_tile(tc)
        nt = (V + tile - 1) // tile
        pm = torch.empty((M, nt), device=DEV); ps = torch.empty((M, nt), device=DEV)
        pi = torch.empty((M, nt), dtype=torch.int32, device=DEV)
        if tc:
            hh, hl = split(h); wh, wl = split(w)
            L.call("navc_vocab_partials_tc", L.TC_BF16X3, L.ptr(hh), L.ptr(hl), K, L.ptr(wh), L.ptr(wl), K, None, M, V, K,
                   L.ptr(pm), L.ptr(ps), L.ptr(pi), None, None, L.stream())
        else:
            L.call("navc_vocab_partials_f32", L.ptr(h), K, L.ptr(w), K, None, M, V, K, L.ptr(pm), L.ptr(ps), L.ptr(pi),
                   None, None, L.stream())
        torch.cuda.synchronize()
        assert (pi[:, 0] == 0).all() and (pm == 0).all()
        assert abs(ps.sum(1)[0].item() - V) < 1e-3


@pytest.mark.parametrize("N,S,D,H,kind,watch", [(7, 9, 128, 8, "NARFormer", 0), (5, 29, 512, 8, "NARFormer", 0),
                                              (6, 15, 128, 8, "ARFormer", 0), (4, 15, 128, 4, "ARFormer", 3),
                                              (3, 40, 128, 8, "SelfMask", 0)])
def test_self_attention(N, S, D, H, kind, watch):
    gen = g(13)
    qkv = torch.randn(N * S, 3 * D, generator=gen)
    toks = torch.randint(1, 50, (N, S), generator=gen)
    for n in range(N):
        toks[n, S - (n % 4):] = 0 if n % 4 else toks[n, S - 1]
    toks[0, 2] = 0  # an interior PAD (predicted <pad>)
    ctx = torch.empty(N * S, D, device=DEV)
    probs = torch.empty(H, N, S, S, device=DEV)
    qkv_d, toks_d = qkv.to(DEV), toks.to(DEV)  # keep device copies alive across the async launch
    L.call("navc_self_attention", L.ptr(qkv_d), 3 * D, L.ptr(toks_d), N, S, D, H, L.MASK_KIND[kind], watch,
           L.ptr(ctx), None, None, L.ptr(probs), L.stream())
    torch.cuda.synchronize()
    dk = D // H
    q, k, v = [t.view(N, S, H, dk).permute(2, 0, 1, 3) for t in qkv.split(D, dim=1)]
    mask = O.self_attention_mask(toks, kind, watch)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(dk)
    sc = sc.masked_fill(mask.unsqueeze(0), O.MASK_FILL)
    p = torch.softmax(sc, -1)
    ref = (p @ v).permute(1, 2, 0, 3).reshape(N * S, D)
    assert (probs.cpu() - p).abs().max().item() < 2e-6
    assert (ctx.cpu() - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("B,group,S,E,D,H", [(3, 1, 9, 16, 128, 8), (4, 3, 15, 12, 128, 8), (2, 6, 29, 120, 512, 8)])
def test_cross_attention(B, group, S, E, D, H):
    gen = g(14)
    N = B * group
    L_ = 2
    q = torch.randn(N * S, D, generator=gen)
    kv = torch.randn(B * E, L_ * 2 * D, generator=gen)
    layer = 1
    ctx = torch.empty(N * S, D, device=DEV)
    probs = torch.empty(H, N, S, E, device=DEV)
    kvd, q_d = kv.to(DEV), q.to(DEV)
    L.call("navc_cross_attention", L.ptr(q_d), D, kvd[:, layer * 2 * D:].data_ptr(), L_ * 2 * D, N, S, E, D, H, group,
           L.ptr(ctx), None, None, L.ptr(probs), L.stream())
    torch.cuda.synchronize()
    dk = D // H
    kl = kv[:, layer * 2 * D: layer * 2 * D + D].view(B, E, H, dk)
    vl = kv[:, layer * 2 * D + D: (layer + 1) * 2 * D].view(B, E, H, dk)
    kl = O.enlarge(kl, group).permute(2, 0, 1, 3)
    vl = O.enlarge(vl, group).permute(2, 0, 1, 3)
    qq = q.view(N, S, H, dk).permute(2, 0, 1, 3)
    p = torch.softmax((qq @ kl.transpose(-1, -2)) / math.sqrt(dk), -1)
    ref = (p @ vl).permute(1, 2, 0, 3).reshape(N * S, D)
    assert (probs.cpu() - p).abs().max().item() < 2e-6
    assert (ctx.cpu() - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("N,S,D,H,kind,watch", [(7, 9, 512, 8, "NARFormer", 0), (768, 28, 512, 8, "NARFormer", 0),
                                              (6, 15, 256, 4, "ARFormer", 0), (9, 15, 128, 2, "ARFormer", 3),
                                              (5, 32, 128, 2, "SelfMask", 0), (3, 1, 64, 1, "NARFormer", 0)])
def test_self_attention_tc(mode, tol, N, S, D, H, kind, watch):
    """tcgen05 self-attention core vs plain torch fp32 on the same (unsplit) inputs."""
    gen = g(15)
    qkv = torch.randn(N * S, 3 * D, generator=gen)
    toks = torch.randint(1, 50, (N, S), generator=gen)
    for n in range(N):
        if n % 4 and S > 4:
            toks[n, S - (n % 4):] = 0
    if S > 4:
        toks[0, 2] = 0
    hi, lo = split(qkv)
    ctx = torch.empty(N * S, D, device=DEV)
    chi = torch.empty(N * S, D, dtype=torch.bfloat16, device=DEV)
    clo = torch.empty(N * S, D, dtype=torch.bfloat16, device=DEV)
    hi_d, lo_d, toks_d = hi.to(DEV), lo.to(DEV), toks.to(DEV)
    L.call("navc_self_attention_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(hi_d), L.ptr(lo_d), 3 * D,
           L.ptr(toks_d), N, S, D, H, L.MASK_KIND[kind], watch, L.ptr(ctx), L.ptr(chi), L.ptr(clo), L.stream())
    torch.cuda.synchronize()
    dk = D // H
    q, k, v = [t.view(N, S, H, dk).permute(2, 0, 1, 3) for t in qkv.split(D, dim=1)]
    mask = O.self_attention_mask(toks, kind, watch)
    sc = ((q @ k.transpose(-1, -2)) / math.sqrt(dk)).masked_fill(mask.unsqueeze(0), O.MASK_FILL)
    ref = (torch.softmax(sc, -1) @ v).permute(1, 2, 0, 3).reshape(N * S, D)
    err = (ctx.cpu() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err
    assert (chi.float().cpu() + clo.float().cpu() - ctx.cpu()).abs().max().item() < 1e-5 * max(4.0, ref.abs().max().item())


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("B,group,S,E,D,H", [(3, 1, 9, 16, 128, 2), (4, 3, 15, 12, 256, 4), (2, 6, 29, 120, 512, 8),
                                             (128, 6, 28, 120, 512, 8), (2, 10, 30, 128, 64, 1)])
def test_cross_attention_tc(mode, tol, B, group, S, E, D, H):
    gen = g(16)
    N = B * group
    L_ = 2
    q = torch.randn(N * S, D, generator=gen)
    kv = torch.randn(B * E, L_ * 2 * D, generator=gen)
    layer = 1
    qh, ql = split(q)
    kh, kl_ = split(kv)
    ctx = torch.empty(N * S, D, device=DEV)
    qh, ql, kh, kl_ = qh.to(DEV), ql.to(DEV), kh.to(DEV), kl_.to(DEV)
    L.call("navc_cross_attention_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(qh), L.ptr(ql), D,
           kh[:, layer * 2 * D:].data_ptr(), kl_[:, layer * 2 * D:].data_ptr(), L_ * 2 * D, N, S, E, D, H, group,
           L.ptr(ctx), None, None, L.stream())
    torch.cuda.synchronize()
    dk = D // H
    kl = kv[:, layer * 2 * D: layer * 2 * D + D].view(B, E, H, dk)
    vl = kv[:, layer * 2 * D + D: (layer + 1) * 2 * D].view(B, E, H, dk)
    kl = O.enlarge(kl, group).permute(2, 0, 1, 3)
    vl = O.enlarge(vl, group).permute(2, 0, 1, 3)
    qq = q.view(N, S, H, dk).permute(2, 0, 1, 3)
    p = torch.softmax((qq @ kl.transpose(-1, -2)) / math.sqrt(dk), -1)
    ref = (p @ vl).permute(1, 2, 0, 3).reshape(N * S, D)
    err = (ctx.cpu() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err


def test_embed_ln_and_layernorm():
    gen = g(15)
    N, S, D, V, group = 12, 11, 128, 60, 3
    word = torch.randn(V, D, generator=gen); word[0].zero_()
    pos = torch.randn(S + 3, D, generator=gen); cat = torch.randn(20, D, generator=gen)
    extra = torch.randn(N // group, D, generator=gen)
    lw = 1 + 0.1 * torch.randn(D, generator=gen); lb = 0.1 * torch.randn(D, generator=gen)
    toks = torch.randint(0, V, (N, S), generator=gen)
    category = torch.randint(0, 20, (N // group, 1), generator=gen)
    out = torch.empty(N * S, D, device=DEV)
    hi = torch.empty(N * S, D, dtype=torch.bfloat16, device=DEV); lo = torch.empty_like(hi)
    dv = [t.to(DEV) for t in (toks, category, word, pos, cat, extra, lw, lb)]  # kept alive
    L.call("navc_embed_ln", L.ptr(dv[0]), L.ptr(dv[1]), L.ptr(dv[2]), L.ptr(dv[3]), L.ptr(dv[4]), L.ptr(dv[5]),
           group, L.ptr(dv[6]), L.ptr(dv[7]), 1e-5, N, S, D, L.ptr(out), L.ptr(hi), L.ptr(lo), L.stream())
    torch.cuda.synchronize()
    e = word[toks] + pos[:S].unsqueeze(0) + O.enlarge(cat[category.squeeze(1)], group).unsqueeze(1) + O.enlarge(extra, group).unsqueeze(1)
    ref = torch.nn.functional.layer_norm(e, (D,), lw, lb, 1e-5).view(N * S, D)
    assert (out.cpu() - ref).abs().max().item() < 5e-6
    assert torch.equal(hi.cpu(), out.cpu().to(torch.bfloat16))
    # layernorm kernel with row mask
    x = torch.randn(50, 512, generator=gen)
    rt = torch.randint(0, 2, (50,), generator=gen)
    lw5 = torch.randn(512, generator=gen); lb5 = torch.randn(512, generator=gen)
    o2 = torch.empty(50, 512, device=DEV)
    dv2 = [t.to(DEV) for t in (x, lw5, lb5, rt)]
    L.call("navc_layernorm", L.ptr(dv2[0]), L.ptr(dv2[1]), L.ptr(dv2[2]), 1e-5, L.ptr(dv2[3]), 50, 512, L.ptr(o2), None, None, L.stream())
    torch.cuda.synchronize()
    ref2 = torch.nn.functional.layer_norm(x, (512,), lw5, lb5, 1e-5) * rt.ne(0).float().unsqueeze(1)
    assert (o2.cpu() - ref2).abs().max().item() < 1e-5


def test_length_beam_canvas_and_select_best():
    gen = g(16)
    B, max_len, lbs = 37, 30, 6
    pred = torch.log_softmax(torch.randn(B, max_len, generator=gen), -1)
    beam = torch.empty(B, lbs, dtype=torch.int32, device=DEV)
    smax = torch.zeros(1, dtype=torch.int32, device=DEV)
    pred_d = pred.to(DEV)
    L.call("navc_length_beam", L.ptr(pred_d), B, max_len, lbs, 0, L.ptr(beam), L.ptr(smax), L.stream())
    ref = O.length_beam(pred, lbs, 0, max_len)
    assert torch.equal(beam.cpu().long(), ref)
    S = int(smax.item())
    assert S == int(ref.max())
    N = B * lbs
    canvas = torch.empty(N, S, dtype=torch.int64, device=DEV); toks = torch.empty_like(canvas)
    probs = torch.empty(N, S, device=DEV)
    L.call("navc_init_canvas", L.ptr(beam), N, S, 4, L.ptr(canvas), L.ptr(toks), L.ptr(probs), L.stream())
    inside = torch.arange(S).view(1, S) < ref.view(-1, 1)
    assert torch.equal(canvas.cpu(), torch.where(inside, 4, 0))
    assert torch.equal(probs.cpu(), (~inside).float())
    lprobs = torch.log(torch.rand(N, S, generator=gen)) * inside
    tk = torch.randint(6, 99, (N, S), generator=gen)
    hyp = torch.empty(B, S, dtype=torch.int64, device=DEV)
    score = torch.empty(N, device=DEV)
    tk_d, lprobs_d = tk.to(DEV), lprobs.to(DEV)
    L.call("navc_select_best", L.ptr(tk_d), L.ptr(lprobs_d), L.ptr(beam), B, lbs, S, 1.35, L.ptr(hyp), L.ptr(score), L.stream())
    sc = lprobs.view(B, lbs, S).sum(-1) / (ref.float() ** 1.35)
    best = sc.max(-1)[1]
    ref_hyp = tk.view(B, lbs, S)[torch.arange(B), best]
    assert torch.equal(hyp.cpu(), ref_hyp)
    assert (score.cpu().view(B, lbs) - sc).abs().max().item() < 1e-5


def _step(N, S, **kw):
    st = L.Step()
    keep = []
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            keep.append(v)
            setattr(st, k, v.data_ptr())
        else:
            setattr(st, k, v)
    L.call("navc_refine_step", st, N, S, L.stream())
    torch.cuda.synchronize()


def test_refine_step_modes_match_oracle_selection():
    gen = g(17)
    N, S, V, tile = 50, 21, 300, 128
    nt = (V + tile - 1) // tile
    lens = torch.randint(4, S + 1, (N,), generator=gen).int()
    inside = torch.arange(S).view(1, S) < lens.view(-1, 1)
    logits = torch.randn(N * S, V, generator=gen)
    logits[5, 4] = 50.0  # force a MASK prediction somewhere inside row 0
    # partials from the logits
    pad = nt * tile - V
    lp = torch.cat([logits, torch.full((N * S, pad), -float("inf"))], 1).view(N * S, nt, tile)
    pm, pidx = lp.max(-1)
    ps = torch.exp(lp - pm.unsqueeze(-1)).sum(-1)
    pidx = (pidx + torch.arange(nt).view(1, nt) * tile).int()
    ref_p, ref_i = torch.softmax(logits, -1).max(-1)
    ref_p, ref_i = ref_p.view(N, S), ref_i.view(N, S)
    ref_i = ref_i.masked_fill(~inside, 0); ref_p = ref_p.masked_fill(~inside, 1.0)
    d = lambda t: t.to(DEV)
    tokens = d(torch.where(inside, 4, 0)); probs = d((~inside).float())
    upd = torch.zeros(N, S, dtype=torch.uint8, device=DEV); canvas = torch.empty(N, S, dtype=torch.int64, device=DEV)
    counters = torch.zeros(2, dtype=torch.int32, device=DEV)
    visual = torch.zeros(N, S, dtype=torch.uint8, device=DEV); masked0 = torch.zeros_like(visual)
    # (1) CT first pass + MASKTOK selection
    _step(N, S, part_max=d(pm), part_sum=d(ps), part_idx=d(pidx), n_tiles=nt, is_ct=1, merge=L.MERGE_ALL,
          select=L.SELECT_MASKTOK, lens=d(lens), tokens=tokens, probs=probs, upd_mask=upd, canvas=canvas,
          counters=counters, visual=visual, masked0=masked0)
    exp_p = ref_p.masked_fill(ref_i.eq(4), 0.0)
    assert torch.equal(tokens.cpu(), ref_i)
    assert (probs.cpu() - exp_p).abs().max().item() < 1e-6
    assert torch.equal(upd.cpu().bool(), ref_i.eq(4))
    assert torch.equal(canvas.cpu(), ref_i)  # masked positions already hold MASK
    assert counters[0].item() == int(ref_i.eq(4).sum()) and counters[1].item() == int(ref_i.eq(4).sum())
    assert torch.equal(visual.cpu().bool(), ref_i.ne(4) & ref_i.ne(0))
    assert torch.equal(masked0.cpu().bool(), ref_i.eq(4) & inside)
    # (2) worst-k selection with a teacher, against the oracle's stable rule
    teacher = torch.rand(N, S, generator=gen)
    ratio = 1.0 - 2 / 5
    _step(N, S, merge=L.MERGE_NONE, select=L.SELECT_WORST, ratio=ratio, lens=d(lens), teacher=d(teacher), tokens=tokens,
          probs=probs, upd_mask=upd, canvas=canvas)
    k = (lens.float() * ratio).long()
    exp_mask = O.k_smallest_mask(exp_p * teacher, k)
    assert torch.equal(upd.cpu().bool(), exp_mask)
    assert torch.equal(canvas.cpu(), ref_i.masked_fill(exp_mask, 4))
    # (3) merge only where masked
    logits2 = torch.randn(N * S, V, generator=gen)
    lp2 = torch.cat([logits2, torch.full((N * S, pad), -float("inf"))], 1).view(N * S, nt, tile)
    pm2, pidx2 = lp2.max(-1); ps2 = torch.exp(lp2 - pm2.unsqueeze(-1)).sum(-1)
    pidx2 = (pidx2 + torch.arange(nt).view(1, nt) * tile).int()
    p2, i2 = torch.softmax(logits2, -1).max(-1)
    p2 = p2.view(N, S).masked_fill(~inside, 1.0); i2 = i2.view(N, S).masked_fill(~inside, 0)
    lprobs = torch.empty(N, S, device=DEV)
    _step(N, S, part_max=d(pm2), part_sum=d(ps2), part_idx=d(pidx2), n_tiles=nt, merge=L.MERGE_MASKED, select=L.SELECT_NONE,
          lens=d(lens), teacher=d(teacher), tokens=tokens, probs=probs, upd_mask=upd, canvas=canvas, lprobs=lprobs)
    exp_tok = torch.where(exp_mask, i2, ref_i); exp_pp = torch.where(exp_mask, p2, exp_p)
    assert torch.equal(tokens.cpu(), exp_tok)
    assert (probs.cpu() - exp_pp).abs().max().item() < 1e-6
    ref_lp = (exp_pp * teacher).log()
    got = lprobs.cpu()
    fin = torch.isfinite(ref_lp)
    assert torch.equal(torch.isfinite(got), fin)
    assert (got[fin] - ref_lp[fin]).abs().max().item() < 1e-5


def test_refine_step_easy_first_and_window():
    gen = g(18)
    N, S, V, tile = 33, 17, 128, 128
    lens = torch.randint(4, S + 1, (N,), generator=gen).int()
    inside = torch.arange(S).view(1, S) < lens.view(-1, 1)
    d = lambda t: t.to(DEV)
    tokens0 = torch.where(inside, 4, 0)
    tokens0[:, 1] = torch.where(inside[:, 1], 77, 0)  # one committed word
    logits = torch.randn(N * S, V, generator=gen)
    pm, pidx = logits.max(-1, keepdim=True); ps = torch.exp(logits - pm).sum(-1, keepdim=True)
    p, i = torch.softmax(logits, -1).max(-1)
    p = p.view(N, S).masked_fill(~inside, 1.0); i = i.view(N, S).masked_fill(~inside, 0)
    for q in (1, 3):
        tokens = d(tokens0.clone()); probs = d((~inside).float())
        upd = torch.zeros(N, S, dtype=torch.uint8, device=DEV); canvas = torch.empty(N, S, dtype=torch.int64, device=DEV)
        counters = torch.zeros(2, dtype=torch.int32, device=DEV)
        _step(N, S, part_max=d(pm), part_sum=d(ps), part_idx=d(pidx.int()), n_tiles=1, merge=L.MERGE_EF, q=q,
              select=L.SELECT_KEEP, lens=d(lens), tokens=tokens, probs=probs, upd_mask=upd, canvas=canvas, counters=counters)
        mask = tokens0.eq(4)
        cand = p.masked_fill(~mask, 0.0)
        commit = O.k_largest_mask(cand, mask.sum(1).clamp(max=q))
        exp_tok = torch.where(commit, i, tokens0)
        assert torch.equal(tokens.cpu(), exp_tok)
        assert torch.equal(canvas.cpu(), exp_tok)
        assert counters[0].item() == int(exp_tok.eq(4).sum())
    # left-to-right window
    given = d(tokens0.eq(4).to(torch.uint8))
    tokens = d(tokens0.clone()); probs = d((~inside).float())
    upd = torch.zeros(N, S, dtype=torch.uint8, device=DEV); canvas = torch.empty(N, S, dtype=torch.int64, device=DEV)
    counters = torch.zeros(2, dtype=torch.int32, device=DEV)
    _step(N, S, merge=L.MERGE_NONE, select=L.SELECT_WINDOW, win_lo=2, win_hi=4, given=given, lens=d(lens), tokens=tokens,
          probs=probs, upd_mask=upd, canvas=canvas, counters=counters)
    m0 = tokens0.eq(4)
    ordinal = m0.long().cumsum(1) - 1
    exp = m0 & (ordinal >= 2) & (ordinal < 4)
    assert torch.equal(upd.cpu().bool(), exp)
    assert counters[1].item() == int(exp.sum())


def test_split_and_log_softmax():
    x = torch.randn(1000, 333, generator=g(19))
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=DEV); lo = torch.empty_like(hi)
    x_d = x.to(DEV)
    L.call("navc_split_bf16", L.ptr(x_d), L.ptr(hi), L.ptr(lo), x.numel(), L.stream())
    h, l = split(x)
    assert torch.equal(hi.cpu(), h) and torch.equal(lo.cpu(), l)
    xd = x.to(DEV).contiguous()
    L.call("navc_log_softmax", L.ptr(xd), L.ptr(xd), 1000, 333, 333, L.stream())
    assert (xd.cpu() - torch.log_softmax(x, -1)).abs().max().item() < 2e-6


# ---------------------------------------------------------------------------------------------------
# packed rows (include/navc.h "packed rows")
# ---------------------------------------------------------------------------------------------------
def _packing(N, S, seed):
    gen = g(seed)
    lens = torch.randint(1, S + 1, (N,), generator=gen).int()
    lens[0] = S
    if N > 3:
        lens[3] = 1
    off = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int32), lens]), 0).int()
    return lens, off


@pytest.mark.parametrize("N,S", [(1, 5), (45, 27), (768, 28), (3000, 30)])
def test_pack_rows(N, S):
    lens, off = _packing(N, S, 41)
    seq_off = torch.empty(N + 1, dtype=torch.int32, device=DEV)
    rowmap = torch.full((N * S,), -1, dtype=torch.int32, device=DEV)
    lens_d = lens.to(DEV)
    L.call("navc_pack_rows", L.ptr(lens_d), N, S, L.ptr(seq_off), L.ptr(rowmap), L.stream())
    assert torch.equal(seq_off.cpu(), off)
    want = torch.cat([n * S + torch.arange(int(lens[n])) for n in range(N)]).int()
    assert torch.equal(rowmap.cpu()[:want.numel()], want)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_linear_device_row_count_and_gather(mode):
    """navc_epilogue_t.m_dev: only the first *m_dev rows are computed (rows beyond keep their old contents);
    navc_gather_rows; navc_vocab_partials_tc_dyn."""
    M, N, K, cnt = 1000, 512, 256, 389
    x = torch.randn(M, K, generator=g(42))
    w = torch.randn(N, K, generator=g(43)) / math.sqrt(K)
    xh, xl = [t.to(DEV) for t in split(x)]
    wh, wl = [t.to(DEV) for t in split(w)]
    ohi = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    olo = torch.full((M, N), 7.0, dtype=torch.bfloat16, device=DEV)
    m_dev = torch.tensor([cnt], dtype=torch.int32, device=DEV)
    ep = L.Epilogue(None, None, None, 0, 0, None, L.ptr(ohi), L.ptr(olo), N, 0, 1, 0, None, None, L.ptr(m_dev))
    L.call("navc_linear_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(xh), L.ptr(xl), K, L.ptr(wh), L.ptr(wl), K,
           M, N, K, ep, L.stream())
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    got = (ohi.float() + olo.float()).cpu()
    tol = 3e-5 if mode == "bf16x3" else 2e-2
    assert (got[:cnt] - ref[:cnt]).abs().max().item() < tol * ref.abs().max().item()
    tile_end = (cnt + 127) // 128 * 128
    assert (ohi[tile_end:].float() == 7.0).all() and (olo[tile_end:].float() == 7.0).all()
    # gather
    rows = torch.randperm(M, generator=g(44))[:300].int().to(DEV)
    count = torch.tensor([257], dtype=torch.int32, device=DEV)
    ghi = torch.zeros((M, K), dtype=torch.bfloat16, device=DEV)
    glo = torch.zeros((M, K), dtype=torch.bfloat16, device=DEV)
    L.call("navc_gather_rows", L.ptr(xh), L.ptr(xl), K, L.ptr(rows), L.ptr(count), M, L.ptr(ghi), L.ptr(glo), L.stream())
    assert torch.equal(ghi[:257], xh[rows[:257].long()]) and torch.equal(glo[:257], xl[rows[:257].long()])
    assert ghi[257:].float().abs().max().item() == 0
    # vocabulary partials over a device-side row count
    V = 777
    wv = torch.randn(V, K, generator=g(45)) / math.sqrt(K)
    vh, vl = [t.to(DEV) for t in split(wv)]
    tile = L._lib.navc_vocab_tile(1)
    nt = (V + tile - 1) // tile
    pm = torch.full((M, nt), 5.0, device=DEV)
    ps = torch.full((M, nt), 5.0, device=DEV)
    pi = torch.full((M, nt), -5, dtype=torch.int32, device=DEV)
    L.call("navc_vocab_partials_tc_dyn", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(xh), L.ptr(xl), K, L.ptr(vh),
           L.ptr(vl), K, None, M, V, K, L.ptr(m_dev), L.ptr(pm), L.ptr(ps), L.ptr(pi), L.stream())
    logits = (x.double() @ wv.double().t())[:cnt]
    m = pm[:cnt].max(1).values.cpu().double()
    assert (m - logits.max(1).values).abs().max().item() < (1e-4 if mode == "bf16x3" else 5e-2)
    if mode == "bf16x3":
        arg = pi[:cnt].gather(1, pm[:cnt].argmax(1, keepdim=True)).squeeze(1).cpu()
        assert (arg == logits.argmax(1)).float().mean().item() > 0.995
    assert (pm[tile_end:] == 5.0).all()


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("kind", ["NARFormer", "ARFormer", "SelfMask"])
@pytest.mark.parametrize("N,S", [(7, 11), (45, 27), (130, 32)])
def test_self_attention_tc_packed(mode, tol, kind, N, S):
    D, H = 512, 8
    lens, off = _packing(N, S, 46)
    R = int(off[-1])
    gen = g(47)
    qkv = torch.randn(N * S, 3 * D, generator=gen)          # packed rows first, garbage (finite) beyond
    toks = torch.zeros(N, S, dtype=torch.int64)
    for n in range(N):
        toks[n, :lens[n]] = torch.randint(1, 50, (int(lens[n]),), generator=gen)
    if lens[1] > 2:
        toks[1, 1] = 0                                       # an interior (predicted) <pad>
    hi, lo = split(qkv)
    chi = torch.zeros(N * S, D, dtype=torch.bfloat16, device=DEV)
    clo = torch.zeros(N * S, D, dtype=torch.bfloat16, device=DEV)
    ctx = torch.zeros(N * S, D, device=DEV)
    hi_d, lo_d, toks_d, off_d = hi.to(DEV), lo.to(DEV), toks.to(DEV), off.to(DEV)   # (kept alive until the kernel has run)
    L.call("navc_self_attention_tc_packed", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(hi_d), L.ptr(lo_d),
           3 * D, L.ptr(toks_d), L.ptr(off_d), N, S, D, H, L.MASK_KIND[kind], 0, L.ptr(ctx), L.ptr(chi), L.ptr(clo),
           L.stream())
    torch.cuda.synchronize()
    dk = D // H
    worst = 0.0
    for n in range(N):
        ln, r0 = int(lens[n]), int(off[n])
        blk = qkv[r0:r0 + ln]
        q, k, v = [t.view(ln, H, dk).permute(1, 0, 2) for t in blk.split(D, dim=1)]
        mask = O.self_attention_mask(toks[n:n + 1, :ln], kind, 0)[0]
        sc = ((q @ k.transpose(-1, -2)) / math.sqrt(dk)).masked_fill(mask.unsqueeze(0), O.MASK_FILL)
        ref = (torch.softmax(sc, -1) @ v).permute(1, 0, 2).reshape(ln, D)
        worst = max(worst, (ctx[r0:r0 + ln].cpu() - ref).abs().max().item() / max(1.0, ref.abs().max().item()))
    assert worst < tol, worst
    assert ctx[R:].abs().max().item() == 0                   # rows beyond the packed count are never written


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("B,group,S,E", [(3, 1, 9, 16), (5, 6, 28, 120), (2, 10, 30, 128)])
def test_cross_attention_tc_packed(mode, tol, B, group, S, E):
    D, H = 512, 8
    N = B * group
    lens, off = _packing(N, S, 48)
    R = int(off[-1])
    gen = g(49)
    q = torch.randn(N * S, D, generator=gen)
    kv = torch.randn(B * E, 2 * D, generator=gen)
    qh, ql = split(q)
    kh, kl_ = split(kv)
    ctx = torch.zeros(N * S, D, device=DEV)
    qh_d, ql_d, kh_d, kl_d, off_d = qh.to(DEV), ql.to(DEV), kh.to(DEV), kl_.to(DEV), off.to(DEV)
    L.call("navc_cross_attention_tc_packed", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(qh_d), L.ptr(ql_d), D,
           L.ptr(kh_d), L.ptr(kl_d), 2 * D, L.ptr(off_d), N, S, E, D, H, group, L.ptr(ctx), None, None, L.stream())
    torch.cuda.synchronize()
    dk = D // H
    worst = 0.0
    for n in range(N):
        ln, r0, b = int(lens[n]), int(off[n]), n // group
        qq = q[r0:r0 + ln].view(ln, H, dk).permute(1, 0, 2)
        kk = kv[b * E:(b + 1) * E, :D].view(E, H, dk).permute(1, 0, 2)
        vv = kv[b * E:(b + 1) * E, D:].view(E, H, dk).permute(1, 0, 2)
        ref = (torch.softmax(qq @ kk.transpose(-1, -2) / math.sqrt(dk), -1) @ vv).permute(1, 0, 2).reshape(ln, D)
        worst = max(worst, (ctx[r0:r0 + ln].cpu() - ref).abs().max().item() / max(1.0, ref.abs().max().item()))
    assert worst < tol, worst
    assert ctx[R:].abs().max().item() == 0


@pytest.mark.parametrize("B,K,V,first", [(7, 5, 10547, 0), (3, 3, 300, 0), (4, 5, 1000, 1), (2, 8, 57, 0), (5, 1, 200, 0)])
def test_beam_topk_matches_torch(B, K, V, first):
    """navc_beam_topk == topk(masked_fill(log_softmax(logits) + beam score, last == EOS, -1e20)) (Beam.py:68-83)."""
    T, pos = 12, 4
    gen = torch.Generator().manual_seed(91)
    ld = V + 3
    logits = torch.randn(B * K, ld, generator=gen) * 3
    scores = -torch.rand(B, K, generator=gen) * 10
    hist = torch.randint(4, 50, (B * K, T), generator=gen)
    if K > 1:
        hist[1, pos] = 3  # an EOS-terminated beam
        hist[K, pos] = 3
    logp = torch.log_softmax(logits[:, :V], -1).view(B, K, V)
    if first:
        lk = logp[:, 0, :]
    else:
        lk = logp + scores.unsqueeze(-1)
        lk = lk.masked_fill(hist[:, pos].view(B, K, 1).eq(3), -1e20).view(B, K * V)
    want_s, want_i = lk.topk(K, dim=1)
    d_logits, d_scores, d_hist = logits.to(DEV), scores.to(DEV), hist.to(DEV)
    bs = torch.empty(B, K, device=DEV)
    bi = torch.empty(B, K, dtype=torch.int64, device=DEV)
    L.call("navc_beam_topk", L.ptr(d_logits), ld, B, K, V, L.ptr(d_scores), L.ptr(d_hist), T, pos, first, L.ptr(bs), L.ptr(bi),
           L.stream())
    assert torch.allclose(bs.cpu(), want_s, atol=2e-5, rtol=1e-6)
    assert torch.equal(bi.cpu(), want_i)


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("M,N,K", [(640, 512, 512), (640, 512, 2048), (640, 2048, 512), (10478, 512, 2048), (10478, 1536, 512),
                                   (4736, 512, 512), (19000, 512, 512), (130, 256, 64), (7, 128, 128), (5000, 2048, 512)])
def test_linear_pair_epilogue_tail_split(mode, tol, M, N, K):
    """Pair epilogue (bf16 hi/lo outputs, bias, gelu, bf16-pair residual, row mask) over grids whose last wave is
    partly filled: those tiles are split along K across the idle SMs and summed by the finishing CTA.  Launched
    three times in a row (the arrival counters must re-arm themselves)."""
    x = torch.randn(M, K, generator=g(50))
    w = torch.randn(N, K, generator=g(51)) / math.sqrt(K)
    b = torch.randn(N, generator=g(52))
    res = torch.randn(M, N, generator=g(53))
    toks = torch.randint(0, 4, (M,), generator=g(54))
    xh, xl = [t.to(DEV) for t in split(x)]
    wh, wl = [t.to(DEV) for t in split(w)]
    rh, rl = [t.to(DEV) for t in split(res)]
    bd, td = b.to(DEV), toks.to(DEV)
    x3 = mode == "bf16x3"
    y = x.double() @ w.double().t() + b.double()
    y = O.activation("gelu_new")(y.float()).double()
    y = (y + (rh.float() + (rl.float() if x3 else 0)).cpu().double()) * toks.ne(0).double().unsqueeze(1)
    prev = L._lib.navc_set_streamk(1)  # opt-in feature (slower than idle SMs in practice): switched on for this test
    try:
        _tail_split_runs(mode, tol, M, N, K, x3, xh, xl, wh, wl, rh, rl, bd, td, y)
    finally:
        L._lib.navc_set_streamk(prev)
    assert L._lib.navc_streamk_error() == 0


def _tail_split_runs(mode, tol, M, N, K, x3, xh, xl, wh, wl, rh, rl, bd, td, y):
    for rep in range(3):
        ohi = torch.full((M, N), 9.0, dtype=torch.bfloat16, device=DEV)
        olo = torch.full((M, N), 9.0, dtype=torch.bfloat16, device=DEV) if x3 else None
        ep = L.Epilogue(L.ptr(bd), None, L.ptr(td), L.ACT["gelu_new"], N, None, L.ptr(ohi), L.ptr(olo), N, 0, 1, 0,
                        L.ptr(rh), L.ptr(rl) if x3 else None, None)
        L.call("navc_linear_tc", L.TC_BF16X3 if x3 else L.TC_BF16, L.ptr(xh), L.ptr(xl) if x3 else None, K, L.ptr(wh),
               L.ptr(wl) if x3 else None, K, M, N, K, ep, L.stream())
        got = (ohi.float() + (olo.float() if x3 else 0)).cpu().double()
        err = (got - y).abs().max().item()
        assert err < tol * max(1.0, y.abs().max().item()), (rep, err)
