"""Training path parity (forward in train mode + hand-written backward) through the C ABI.

Kernel level: every backward kernel against torch autograd of the same op on the CPU (fp32;
tolerance 1e-5 of the output scale -- accumulation order only).  Model level: loss and the
gradient of EVERY parameter against the oracle (``oracle/navc_oracle.py`` under torch autograd, which
``tests/test_oracle_vs_reference.py`` pins to the reference's own ``loss.backward()``) with dropout
disabled and BatchNorm in train mode; tolerances: fp32 mode 2e-4, bf16x3
mode 5e-4 of each gradient's max magnitude.  Dropout itself (not bit-reproducible against torch's
Philox stream) is checked for keep-rate, scaling and forward/backward mask agreement."""
import math

import pytest
import torch
import torch.nn.functional as F

import cases
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


_KEEP = []


def P(t):
    """Device pointer of a (copied) tensor that stays alive for the rest of the test module: a
    temporary passed as P(x) would be freed -- and its block re-used by the next
    temporary -- before the kernel even launches."""
    t = t.detach().to(DEV).contiguous()
    _KEEP.append(t)
    if len(_KEEP) > 64:
        torch.cuda.synchronize()
        del _KEEP[:32]
    return L.ptr(t)


def zero_atol(name, ref_grad_of, base=2e-7):
    """Absolute slack for gradients that are analytically zero: softmax is invariant to a key bias, so what a kernel
    returns for `*.key.bias` is rounding noise of dS (fp32 CUDA cores: ~1e-8; split-bf16 tcgen05 operands: 2^-17 relative
    per element) -- judged against the scale of the key WEIGHT gradient of the same block."""
    if name.endswith("key.bias"):
        return base + 1e-4 * ref_grad_of(name.replace("key.bias", "key.weight")).abs().max().item()
    return base


def close(a, b, tol, what="", atol=0.0):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = max(b.abs().max().item(), 1e-6)
    err = (a - b).abs().max().item()
    assert err < tol * scale + atol, "%s: abs err %.3e, rel %.3e (scale %.3e)" % (what, err, err / scale, scale)


# ---------------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------------
def test_drop_add_masks_agree_and_rates():
    M, D = 512, 256
    y = torch.ones(M, D, device=DEV)
    res = torch.full((M, D), 3.0, device=DEV)
    toks = torch.randint(0, 4, (M,), generator=g(1)).to(DEV)
    out = torch.empty(M, D, device=DEV)
    hi = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
    lo = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
    s1, s2 = 12345, 2 ** 63 + 99
    L.call("navc_drop_add", L.ptr(y), L.ptr(res), s1, 0.5, s2, 0.25, L.ptr(toks), M, D, L.ptr(out), L.ptr(hi), L.ptr(lo), L.stream())
    d_y = torch.empty(M, D, device=DEV)
    d_res = torch.empty(M, D, device=DEV)
    L.call("navc_drop_add_bwd", P(torch.ones(M, D)), s1, 0.5, s2, 0.25, L.ptr(toks), M, D, L.ptr(d_y), L.ptr(d_res), L.stream())
    torch.cuda.synchronize()
    live = toks.ne(0).unsqueeze(1).expand(-1, D)
    assert out[~live].abs().max().item() == 0 and d_y[~live].abs().max().item() == 0
    # out = m2*(m1*1 + 3) with m1 in {0,2}, m2 in {0,4/3}; d_res = m2, d_y = m1*m2
    m2 = d_res[live]
    m1m2 = d_y[live]
    assert all(v == 0.0 or abs(v - 4 / 3) < 1e-6 for v in torch.unique(m2).tolist())
    keep2 = (m2 > 0).float().mean().item()
    keep1 = (m1m2[m2 > 0] > 0).float().mean().item()
    assert abs(keep2 - 0.75) < 0.01 and abs(keep1 - 0.5) < 0.01
    m1 = torch.where(m2 > 0, m1m2 / m2.clamp_min(1e-9), torch.zeros_like(m2))
    assert torch.allclose(out[live], m2 * (m1 + 3.0), atol=1e-6)
    assert torch.allclose(hi.float() + lo.float(), out, atol=1e-4)
    # different seeds -> different masks; p = 0 -> identity
    out2 = torch.empty(M, D, device=DEV)
    L.call("navc_drop_add", L.ptr(y), L.ptr(res), s1 + 1, 0.5, s2, 0.25, L.ptr(toks), M, D, L.ptr(out2), None, None, L.stream())
    assert not torch.equal(out, out2)
    L.call("navc_drop_add", L.ptr(y), L.ptr(res), 0, 0.0, 0, 0.0, None, M, D, L.ptr(out2), None, None, L.stream())
    assert torch.equal(out2, y + res)


@pytest.mark.parametrize("act", ["none", "gelu_new", "gelu", "relu", "swish"])
def test_act_drop_backward(act):
    n = 4099
    u = (2.0 * torch.randn(n, generator=g(2))).requires_grad_(True)
    dout = torch.randn(n, generator=g(3))
    ref = O.activation(act)(u) if act != "none" else u * 1.0
    ref.backward(dout)
    ud, dd = u.detach().to(DEV), dout.to(DEV)
    out = torch.empty(n, device=DEV)
    du = torch.empty(n, device=DEV)
    L.call("navc_act_drop", L.ptr(ud), L.ACT[act], 0, 0.0, n, L.ptr(out), None, None, L.stream())
    L.call("navc_act_drop_bwd", L.ptr(dd), L.ptr(ud), L.ACT[act], 0, 0.0, n, L.ptr(du), L.stream())
    close(out, ref, 2e-6, "act fwd")
    close(du, u.grad, 1e-5, "act bwd")
    # with dropout: backward mask == forward mask
    L.call("navc_act_drop", P(torch.ones(n)), 0, 77, 0.5, n, L.ptr(out), None, None, L.stream())
    L.call("navc_act_drop_bwd", P(torch.ones(n)), L.ptr(ud), 0, 77, 0.5, n, L.ptr(du), L.stream())
    assert torch.equal(out, du) and abs((out > 0).float().mean().item() - 0.5) < 0.03


@pytest.mark.parametrize("M,N,ld", [(100, 70, 72), (64, 128, 128), (333, 30, 64), (1000, 517, 520)])
def test_transpose_pack(M, N, ld):
    x = torch.zeros(M, ld)
    x[:, :N] = torch.randn(M, N, generator=g(4))
    xd = x.to(DEV)
    Mp = (M + 63) // 64 * 64
    hi = torch.full((M, ld), 7.0, dtype=torch.bfloat16, device=DEV)
    lo = torch.full((M, ld), 7.0, dtype=torch.bfloat16, device=DEV)
    t32 = torch.full((N, Mp), 7.0, device=DEV)
    thi = torch.full((N, Mp), 7.0, dtype=torch.bfloat16, device=DEV)
    tlo = torch.full((N, Mp), 7.0, dtype=torch.bfloat16, device=DEV)
    cs = torch.zeros(N, device=DEV)
    L.call("navc_transpose_pack", L.ptr(xd), M, N, ld, L.ptr(hi), L.ptr(lo), ld, L.ptr(t32), L.ptr(thi), L.ptr(tlo), Mp, L.ptr(cs), L.stream())
    torch.cuda.synchronize()
    assert torch.equal(t32[:, :M].cpu(), x[:, :N].t())
    assert t32[:, M:].abs().max().item() == 0 if Mp > M else True
    assert torch.allclose((hi.float() + lo.float()).cpu(), x, atol=1e-5) and hi[:, N:].float().abs().max().item() == 0 if ld > N else True
    assert torch.allclose((thi.float() + tlo.float())[:, :M].cpu(), x[:, :N].t(), atol=1e-5)
    close(cs, x[:, :N].sum(0), 1e-5, "colsum")


@pytest.mark.parametrize("M,N,ld", [(70, 64, 64), (333, 520, 576), (4195, 2048, 2048), (130, 30, 64), (1, 10547, 10560), (77, 10, 16)])
@pytest.mark.parametrize("with_lo", [True, False])
def test_split_colsum_fast_path(M, N, ld, with_lo):
    """navc_transpose_pack without transposed outputs (what every gradient GEMM of the tensor-core path asks for): the
    vectorised split + column-sum kernel; pad columns N..ld-1 come out as zeros, rows beyond M are not touched."""
    x = torch.zeros(M, ld)
    x[:, :N] = torch.randn(M, N, generator=g(4))
    x[:, N:] = 123.0      # garbage in the source's pad columns must not leak
    xd = x.to(DEV)
    hi = torch.full((M + 3, ld), 7.0, dtype=torch.bfloat16, device=DEV)
    lo = torch.full((M + 3, ld), 7.0, dtype=torch.bfloat16, device=DEV) if with_lo else None
    cs = torch.zeros(N, device=DEV)
    L.call("navc_transpose_pack", L.ptr(xd), M, N, ld, L.ptr(hi), L.ptr(lo), ld, None, None, None, 0, L.ptr(cs), L.stream())
    want = x.clone()
    want[:, N:] = 0.0
    got = hi[:M].float() + (lo[:M].float() if with_lo else 0.0)
    assert torch.allclose(got.cpu(), want, atol=1e-5 if with_lo else 4e-2)
    assert torch.equal(hi[:M].cpu(), want.to(torch.bfloat16))
    assert torch.all(hi[M:].float() == 7.0)
    close(cs, x[:, :N].sum(0), 1e-5, "colsum")


@pytest.mark.parametrize("mode", ["f32", "bf16x3"])
def test_gemm_split_k_accumulate(mode):
    M, N, K = 300, 520, 4096 + 64
    x = torch.randn(M, K, generator=g(5)).to(DEV)
    w = (torch.randn(N, K, generator=g(6)) / math.sqrt(K)).to(DEV)
    base = torch.randn(M, N, generator=g(7)).to(DEV)
    out = base.clone()
    ref = base.double() + x.double() @ w.double().t()
    ep = L.Epilogue(None, None, None, 0, 0, L.ptr(out), None, None, N, 0, 7, 1)
    if mode == "f32":
        L.call("navc_linear_f32", L.ptr(x), K, L.ptr(w), K, M, N, K, ep, L.stream())
    else:
        xh, xl = x.to(torch.bfloat16), (x - x.to(torch.bfloat16).float()).to(torch.bfloat16)
        wh, wl = w.to(torch.bfloat16), (w - w.to(torch.bfloat16).float()).to(torch.bfloat16)
        L.call("navc_linear_tc", L.TC_BF16X3, L.ptr(xh), L.ptr(xl), K, L.ptr(wh), L.ptr(wl), K, M, N, K, ep, L.stream())
    close(out, ref, 3e-5, "split-k accumulate")


def test_gemm_tc_k_tail():
    """K not a multiple of 64: the tail of the last k-block is zero-filled by TMA."""
    M, N, K = 200, 256, 200
    x = torch.randn(M, K, generator=g(8)).to(DEV)
    w = torch.randn(N, K, generator=g(9)).to(DEV)
    out = torch.empty(M, N, device=DEV)
    xh, xl = x.to(torch.bfloat16), (x - x.to(torch.bfloat16).float()).to(torch.bfloat16)
    wh, wl = w.to(torch.bfloat16), (w - w.to(torch.bfloat16).float()).to(torch.bfloat16)
    ep = L.Epilogue(None, None, None, 0, 0, L.ptr(out), None, None, N, 0, 1, 0)
    L.call("navc_linear_tc", L.TC_BF16X3, L.ptr(xh), L.ptr(xl), K, L.ptr(wh), L.ptr(wl), K, M, N, K, ep, L.stream())
    close(out, x.double() @ w.double().t(), 3e-5, "k tail")


@pytest.mark.parametrize("gate", [1, 0])
def test_highway_train_and_backward(gate):
    BF, D = 77, 64
    x = torch.randn(BF, D, generator=g(10)).requires_grad_(True)
    yg = torch.randn(BF, (2 if gate else 1) * D, generator=g(11)).requires_grad_(True)
    y = torch.tanh(yg[:, :D])
    if gate:
        gt = torch.sigmoid(yg[:, D:])
        ref = gt * x + (1 - gt) * y
    else:
        ref = x + y
    d_o = torch.randn(BF, D, generator=g(12))
    ref.backward(d_o)
    xd, ygd = x.detach().to(DEV), yg.detach().to(DEV)
    o = torch.empty(BF, D, device=DEV)
    L.call("navc_highway_fwd_train", L.ptr(xd), L.ptr(ygd), gate, BF, D, 0, 0.0, L.ptr(o), L.stream())
    dx = torch.empty(BF, D, device=DEV)
    dyg = torch.empty_like(ygd)
    L.call("navc_highway_bwd", P(d_o), L.ptr(xd), L.ptr(ygd), gate, BF, D, 0, 0.0, L.ptr(dx), L.ptr(dyg), L.stream())
    close(o, ref, 2e-6, "highway fwd")
    close(dx, x.grad, 1e-5, "highway dx")
    close(dyg, yg.grad, 1e-5, "highway dyg")


@pytest.mark.parametrize("with_bn", [True, False])
def test_batchnorm_train_forward_backward(with_bn):
    B, F_, D, E, slot, nm = 5, 6, 96, 12, 1, 2
    o = (1.5 * torch.randn(B * F_, D, generator=g(13)) + 0.3).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(D, generator=g(14))).requires_grad_(True)
    b = (0.1 * torch.randn(D, generator=g(15))).requires_grad_(True)
    rm, rv = 0.1 * torch.randn(D, generator=g(16)), 0.5 + torch.rand(D, generator=g(17))
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y = F.batch_norm(o, rm_ref, rv_ref, w, b, True, 0.1, 1e-5) if with_bn else o
    hid = o.view(B, F_, D).mean(1) / nm
    d_enc = torch.randn(B, E, D, generator=g(18))
    d_hid = torch.randn(B, D, generator=g(19))
    ((y.view(B, F_, D) * d_enc[:, slot * F_:(slot + 1) * F_]).sum() + (hid * d_hid).sum()).backward()
    od = o.detach().to(DEV)
    mean = torch.empty(D, device=DEV)
    var = torch.empty(D, device=DEV)
    rmd, rvd = rm.to(DEV), rv.to(DEV)
    if with_bn:
        L.call("navc_bn_stats", L.ptr(od), B * F_, D, L.ptr(mean), L.ptr(var), 0.1, L.ptr(rmd), L.ptr(rvd), L.stream())
        close(rmd, rm_ref, 1e-5, "running_mean")
        close(rvd, rv_ref, 1e-5, "running_var")
    enc = torch.zeros(B, E, D, device=DEV)
    ehid = torch.zeros(B, D, device=DEV)
    wd, bd = w.detach().to(DEV), b.detach().to(DEV)
    L.call("navc_bn_apply_concat", L.ptr(od), L.ptr(mean) if with_bn else None, L.ptr(var) if with_bn else None,
           L.ptr(wd) if with_bn else None, L.ptr(bd) if with_bn else None, 1e-5, B, F_, D, E, slot, nm, 0, L.ptr(ehid),
           L.ptr(enc), None, None, L.stream())
    close(enc[:, slot * F_:(slot + 1) * F_], y.view(B, F_, D), 1e-5, "bn fwd")
    close(ehid, hid, 1e-5, "enc_hidden")
    d_w, d_b, d_o = torch.empty(D, device=DEV), torch.empty(D, device=DEV), torch.empty(B * F_, D, device=DEV)
    L.call("navc_bn_bwd", P(d_enc), P(d_hid), L.ptr(od), L.ptr(mean) if with_bn else None,
           L.ptr(var) if with_bn else None, L.ptr(wd) if with_bn else None, 1e-5, B, F_, D, E, slot, nm,
           L.ptr(d_w) if with_bn else None, L.ptr(d_b) if with_bn else None, L.ptr(d_o), L.stream())
    close(d_o, o.grad, 2e-5, "bn d_o")
    if with_bn:
        close(d_w, w.grad, 2e-5, "bn d_w")
        close(d_b, b.grad, 2e-5, "bn d_b")


def test_mean_bwd_and_log_softmax_bwd():
    B, E, D = 4, 7, 64
    dm = torch.randn(B, D, generator=g(20))
    base = torch.randn(B, E, D, generator=g(21))
    acc = base.clone().to(DEV)
    L.call("navc_mean_bwd", P(dm), B, E, D, L.ptr(acc), L.stream())
    close(acc, base + dm.unsqueeze(1) / E, 1e-6, "mean bwd")
    M, V, Vp = 37, 301, 320
    logits = torch.randn(M, V, generator=g(22)).requires_grad_(True)
    lp = torch.log_softmax(logits, -1)
    gg = torch.randn(M, V, generator=g(23)) * (torch.rand(M, V, generator=g(24)) < 0.05).float()
    lp.backward(gg)
    out = torch.full((M, Vp), 9.0, device=DEV)
    L.call("navc_log_softmax_bwd", P(gg), P(lp), M, V, V, L.ptr(out), Vp, L.stream())
    close(out[:, :V], logits.grad, 1e-5, "log_softmax bwd")
    assert out[:, V:].abs().max().item() == 0


def test_layernorm_backward():
    M, D = 70, 128
    x = torch.randn(M, D, generator=g(25)).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(D, generator=g(26))).requires_grad_(True)
    b = torch.zeros(D, requires_grad=True)
    toks = torch.randint(0, 3, (M,), generator=g(27))
    y = F.layer_norm(x, (D,), w, b, 1e-5) * toks.ne(0).float().unsqueeze(1)
    dy = torch.randn(M, D, generator=g(28))
    y.backward(dy)
    dx = torch.empty(M, D, device=DEV)
    dw, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    L.call("navc_layernorm_bwd", P(dy), P(x), P(w), 1e-5,
           P(toks), M, D, L.ptr(dx), L.ptr(dw), L.ptr(db), L.stream())
    close(dx, x.grad, 2e-5, "ln dx")
    close(dw, w.grad, 2e-5, "ln dw")
    close(db, b.grad, 2e-5, "ln db")


def test_embed_ln_backward():
    N, S, D, V, C, group = 6, 9, 128, 50, 5, 2
    word = torch.randn(V, D, generator=g(29)).requires_grad_(True)
    pos = torch.randn(S + 3, D, generator=g(30)).requires_grad_(True)
    cat = torch.randn(C, D, generator=g(31)).requires_grad_(True)
    extra = torch.randn(N // group, D, generator=g(32)).requires_grad_(True)
    lw = (1 + 0.1 * torch.randn(D, generator=g(33))).requires_grad_(True)
    lb = (0.1 * torch.randn(D, generator=g(34))).requires_grad_(True)
    toks = torch.randint(0, V, (N, S), generator=g(35))
    toks[:, -2:] = 0
    cats = torch.randint(0, C, (N // group, 1), generator=g(36))
    idx = torch.arange(N) // group
    e = F.embedding(toks, word, padding_idx=0) + pos[:S].unsqueeze(0) + cat[cats[idx, 0]].unsqueeze(1) + extra[idx].unsqueeze(1)
    y = F.layer_norm(e, (D,), lw, lb, 1e-5)
    dy = torch.randn(N, S, D, generator=g(37))
    y.backward(dy)
    outs = [torch.zeros_like(t, device=DEV) for t in (word, pos, cat, extra, lw, lb)]
    L.call("navc_embed_ln_bwd", P(dy), P(toks), P(cats), P(word), P(pos),
           P(cat), P(extra), group, P(lw), P(lb), 1e-5, N, S, D, *[L.ptr(o) for o in outs], L.stream())
    for o, r, name in zip(outs, (word, pos, cat, extra, lw, lb), ("word", "pos", "cat", "extra", "ln_w", "ln_b")):
        close(o, r.grad, 3e-5, "embed " + name)
    assert outs[0][0].abs().max().item() == 0  # padding_idx row


def _mha_ref(q, k, v, mask, H):
    n, sq, d = q.shape
    sk = k.shape[1]
    dk = d // H
    qh = q.view(n, sq, H, dk).permute(2, 0, 1, 3)
    kh = k.view(n, sk, H, dk).permute(2, 0, 1, 3)
    vh = v.view(n, sk, H, dk).permute(2, 0, 1, 3)
    sc = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(dk)
    if mask is not None:
        sc = sc.masked_fill(mask.unsqueeze(0), O.MASK_FILL)
    p = torch.softmax(sc, -1)
    return torch.matmul(p, vh).permute(1, 2, 0, 3).contiguous().view(n, sq, d)


@pytest.mark.parametrize("kind", ["NARFormer", "ARFormer", "SelfMask"])
@pytest.mark.parametrize("D,H,S", [(128, 8, 11), (512, 8, 30), (64, 2, 7)])
def test_self_attention_backward(kind, D, H, S):
    N = 5
    qkv = torch.randn(N, S, 3 * D, generator=g(38)).requires_grad_(True)
    toks = torch.randint(1, 9, (N, S), generator=g(39))
    for n in range(N):
        toks[n, S - n % 4:] = 0 if n % 4 else toks[n, S - 1]
    mask = O.self_attention_mask(toks, kind, 0)
    ctx = _mha_ref(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], mask, H)
    d_ctx = torch.randn(N, S, D, generator=g(40))
    ctx.backward(d_ctx)
    d_qkv = torch.empty(N * S, 3 * D, device=DEV)
    L.call("navc_self_attention_bwd", P(qkv), 3 * D, P(toks), N, S, D, H, L.MASK_KIND[kind], 0,
           P(d_ctx), L.ptr(d_qkv), L.stream())
    close(d_qkv.view(N, S, 3 * D), qkv.grad, 2e-5, "self attn bwd")


@pytest.mark.parametrize("D,H,S,E,group", [(128, 8, 9, 12, 1), (512, 8, 30, 120, 1), (128, 4, 7, 16, 3), (128, 8, 9, 13, 6)])
def test_cross_attention_backward(D, H, S, E, group):
    B = 4
    N = B * group
    q = torch.randn(N, S, D, generator=g(41)).requires_grad_(True)
    kv = torch.randn(B, E, 2 * D, generator=g(42)).requires_grad_(True)
    kve = kv.unsqueeze(1).expand(-1, group, -1, -1).reshape(N, E, 2 * D)
    ctx = _mha_ref(q, kve[..., :D], kve[..., D:], None, H)
    d_ctx = torch.randn(N, S, D, generator=g(43))
    ctx.backward(d_ctx)
    d_q = torch.empty(N * S, D, device=DEV)
    d_kv = torch.empty(B * E, 2 * D, device=DEV)
    L.call("navc_cross_attention_bwd", P(q), D, P(kv), 2 * D, N, S, E, D, H, group,
           P(d_ctx), L.ptr(d_q), D, L.ptr(d_kv), 2 * D, L.stream())
    close(d_q.view(N, S, D), q.grad, 2e-5, "cross attn dq")
    close(d_kv.view(B, E, 2 * D), kv.grad, 2e-5, "cross attn dkv")


# ---------------------------------------------------------------------------------------------------
# packed rows (training): only the real positions are rows; same numbers as the padded kernels
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,N,K", [(200, 512, 512), (4195, 2048, 512), (333, 10547, 512), (77, 1536, 512), (130, 30, 512),
                                   (64, 512, 2048), (300, 136, 72)])
def test_dgrad_tc_matches_fp32(mode, M, N, K):
    """dX = dY W (+ residual) from row-major bf16 weights; N need not be a multiple of 8 (padded dY columns)."""
    ld = (N + 63) // 64 * 64
    dy = torch.zeros(M, ld)
    dy[:, :N] = torch.randn(M, N, generator=g(80))
    w = torch.randn(N, K, generator=g(81)) * 0.1
    res = torch.randn(M, K, generator=g(82))
    want = dy[:, :N].double() @ w.double() + res.double()
    dyd, wd = dy.to(DEV), w.to(DEV)
    hi = lambda t: t.to(torch.bfloat16)
    lo = lambda t: (t - hi(t).float()).to(torch.bfloat16)
    dy_hi, dy_lo, w_hi, w_lo = hi(dyd), lo(dyd), hi(wd), lo(wd)
    x3 = mode == "bf16x3"
    out = torch.full((M, K), float("nan"), device=DEV)
    resd = res.to(DEV)
    ep = L.Epilogue(None, L.ptr(resd), None, 0, K, L.ptr(out), None, None, K, 0, 1, 0)
    L.call("navc_dgrad_tc", L.TC_BF16X3 if x3 else L.TC_BF16, L.ptr(dy_hi), L.ptr(dy_lo) if x3 else None, ld, L.ptr(w_hi),
           L.ptr(w_lo) if x3 else None, K, M, N, K, ep, L.stream())
    # bf16x3 carries ~16 mantissa bits per operand: ~2e-5 relative per product, accumulated over N terms
    close(out, want, (2e-5 if N < 4096 else 1e-4) if x3 else 1.5e-2, "dgrad %s" % mode)


def _pack(lens, S):
    off = [0]
    for l in lens:
        off.append(off[-1] + l)
    rowmap = [n * S + s for n, l in enumerate(lens) for s in range(l)]
    return torch.tensor(off, dtype=torch.int32), torch.tensor(rowmap, dtype=torch.int64)


@pytest.mark.parametrize("kind", ["NARFormer", "ARFormer", "SelfMask"])
@pytest.mark.parametrize("D,H,S", [(128, 2, 11), (512, 8, 30), (64, 2, 7)])
def test_self_attention_backward_packed(kind, D, H, S):
    lens = [S, 1, S - 3, 5, 2, S - 1]
    N = len(lens)
    seq_off, rowmap = _pack(lens, S)
    Rp = int(seq_off[-1])
    toks = torch.randint(1, 9, (N, S), generator=g(60))
    for n, l in enumerate(lens):
        toks[n, l:] = 0
    qkv = torch.randn(Rp, 3 * D, generator=g(61)).requires_grad_(True)
    d_ctx = torch.randn(Rp, D, generator=g(62))
    for n, l in enumerate(lens):  # reference: every sequence on its own, only its real positions
        r0 = int(seq_off[n])
        x = qkv[r0:r0 + l].unsqueeze(0)
        mask = O.self_attention_mask(toks[n:n + 1, :l], kind, 0)
        ctx = _mha_ref(x[..., :D], x[..., D:2 * D], x[..., 2 * D:], mask, H)
        ctx.backward(d_ctx[r0:r0 + l].unsqueeze(0))
    d_qkv = torch.full((Rp, 3 * D), float("nan"), device=DEV)
    L.call("navc_self_attention_bwd_packed", P(qkv), 3 * D, P(toks), P(seq_off), N, S, D, H, L.MASK_KIND[kind], 0,
           P(d_ctx), L.ptr(d_qkv), L.stream())
    close(d_qkv, qkv.grad, 2e-5, "packed self attn bwd")


@pytest.mark.parametrize("D,H,S,E", [(128, 2, 9, 12), (512, 8, 30, 120), (128, 4, 7, 16)])
def test_cross_attention_backward_packed(D, H, S, E):
    lens = [S, 1, S - 3, 5, 2]
    N = len(lens)
    seq_off, _ = _pack(lens, S)
    Rp = int(seq_off[-1])
    q = torch.randn(Rp, D, generator=g(63)).requires_grad_(True)
    kv = torch.randn(N, E, 2 * D, generator=g(64)).requires_grad_(True)
    d_ctx = torch.randn(Rp, D, generator=g(65))
    for n, l in enumerate(lens):
        r0 = int(seq_off[n])
        ctx = _mha_ref(q[r0:r0 + l].unsqueeze(0), kv[n:n + 1, :, :D], kv[n:n + 1, :, D:], None, H)
        ctx.backward(d_ctx[r0:r0 + l].unsqueeze(0))
    d_q = torch.full((Rp, D), float("nan"), device=DEV)
    d_kv = torch.full((N * E, 2 * D), float("nan"), device=DEV)
    L.call("navc_cross_attention_bwd_packed", P(q), D, P(kv), 2 * D, P(seq_off), N, S, E, D, H, P(d_ctx), L.ptr(d_q), D,
           L.ptr(d_kv), 2 * D, L.stream())
    close(d_q, q.grad, 2e-5, "packed cross attn dq")
    close(d_kv.view(N, E, 2 * D), kv.grad, 2e-5, "packed cross attn dkv")


@pytest.mark.parametrize("with_ctx", [True, False])
@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("kind", ["NARFormer", "ARFormer", "SelfMask"])
@pytest.mark.parametrize("D,H,S", [(128, 2, 11), (512, 8, 30), (64, 1, 32)])
def test_self_attention_backward_tc(mode, kind, D, H, S, with_ctx):
    """csrc/attention_bwd_tc.cu (tcgen05, dk = 64) against torch autograd, sequence by sequence; rows beyond the
    packed count are NaN-poisoned and must neither be read into a result nor written."""
    lens = [min(l, S) for l in (S, 1, S - 3, 5, 2, S - 1, 16, 17)]
    N = len(lens)
    seq_off, rowmap = _pack(lens, S)
    Rp = int(seq_off[-1])
    toks = torch.randint(1, 9, (N, S), generator=g(60))
    for n, l in enumerate(lens):
        toks[n, l:] = 0
    qkv = torch.randn(Rp, 3 * D, generator=g(61)).requires_grad_(True)
    d_ctx = torch.randn(Rp, D, generator=g(62))
    ctx_rows = []
    for n, l in enumerate(lens):
        r0 = int(seq_off[n])
        x = qkv[r0:r0 + l].unsqueeze(0)
        mask = O.self_attention_mask(toks[n:n + 1, :l], kind, 0)
        ctx = _mha_ref(x[..., :D], x[..., D:2 * D], x[..., 2 * D:], mask, H)
        ctx.backward(d_ctx[r0:r0 + l].unsqueeze(0))
        ctx_rows.append(ctx.detach()[0])
    pad = 40   # poisoned rows behind the packed rows
    ctx_d = torch.full((Rp + pad, D), float("nan"), device=DEV)
    ctx_d[:Rp] = torch.cat(ctx_rows).to(DEV)
    qkv_d = torch.full((Rp + pad, 3 * D), float("nan"), device=DEV)
    qkv_d[:Rp] = qkv.detach().to(DEV)
    dctx_d = torch.full((Rp + pad, D), float("nan"), device=DEV)
    dctx_d[:Rp] = d_ctx.to(DEV)
    d_qkv = torch.full((Rp + pad, 3 * D), 7.0, device=DEV)
    L.call("navc_self_attention_bwd_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(qkv_d), 3 * D, P(seq_off), N, S, D, H,
           L.MASK_KIND[kind], 0, L.ptr(dctx_d), L.ptr(ctx_d) if with_ctx else None, L.ptr(d_qkv), L.stream())
    close(d_qkv[:Rp], qkv.grad, 3e-5 if mode == "bf16x3" else 2e-2, "tc self attn bwd %s" % mode)
    assert torch.all(d_qkv[Rp:] == 7.0)


@pytest.mark.parametrize("with_ctx", [True, False])
@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("D,H,S,E", [(128, 2, 9, 12), (512, 8, 30, 120), (64, 1, 32, 128)])
def test_cross_attention_backward_tc(mode, D, H, S, E, with_ctx):
    lens = [min(l, S) for l in (S, 1, S - 3, 5, 2, 16, 17)]
    N = len(lens)
    seq_off, _ = _pack(lens, S)
    Rp = int(seq_off[-1])
    q = torch.randn(Rp, D, generator=g(63)).requires_grad_(True)
    kv = torch.randn(N, E, 2 * D, generator=g(64)).requires_grad_(True)
    d_ctx = torch.randn(Rp, D, generator=g(65))
    ctx_rows = []
    for n, l in enumerate(lens):
        r0 = int(seq_off[n])
        ctx = _mha_ref(q[r0:r0 + l].unsqueeze(0), kv[n:n + 1, :, :D], kv[n:n + 1, :, D:], None, H)
        ctx.backward(d_ctx[r0:r0 + l].unsqueeze(0))
        ctx_rows.append(ctx.detach()[0])
    pad = 40
    ctx_d = torch.full((Rp + pad, D), float("nan"), device=DEV)
    ctx_d[:Rp] = torch.cat(ctx_rows).to(DEV)
    q_d = torch.full((Rp + pad, D), float("nan"), device=DEV)
    q_d[:Rp] = q.detach().to(DEV)
    dctx_d = torch.full((Rp + pad, D), float("nan"), device=DEV)
    dctx_d[:Rp] = d_ctx.to(DEV)
    d_q = torch.full((Rp + pad, D), 7.0, device=DEV)
    d_kv = torch.full((N * E, 2 * D), float("nan"), device=DEV)
    L.call("navc_cross_attention_bwd_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(q_d), D, P(kv), 2 * D, P(seq_off),
           N, S, E, D, H, L.ptr(dctx_d), L.ptr(ctx_d) if with_ctx else None, L.ptr(d_q), D, L.ptr(d_kv), 2 * D, L.stream())
    tol = 3e-5 if mode == "bf16x3" else 2e-2
    close(d_q[:Rp], q.grad, tol, "tc cross attn dq %s" % mode)
    close(d_kv.view(N, E, 2 * D), kv.grad, tol, "tc cross attn dkv %s" % mode)
    assert torch.all(d_q[Rp:] == 7.0)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_cross_attention_backward_tc_split_output(mode):
    """navc_cross_attention_bwd_tc_split: dK / dV leave as bf16 hi (/ lo) rows inside a wider [rows, L * 2D] operand, their
    column sums accumulate into the bias-gradient slice; same numbers as the fp32 result of the plain kernel."""
    D, H, S, E, Lw = 128, 2, 9, 12, 3     # operand of 3 "layers": this launch fills columns [2D, 4D) of it
    lens = [S, 1, S - 3, 5, 2, 7]
    N = len(lens)
    seq_off, _ = _pack(lens, S)
    Rp = int(seq_off[-1])
    q = torch.randn(Rp, D, generator=g(63)).to(DEV)
    kv = torch.randn(N * E, 2 * D, generator=g(64)).to(DEV)
    d_ctx = torch.randn(Rp, D, generator=g(65)).to(DEV)
    x3 = mode == "bf16x3"
    md = L.TC_BF16X3 if x3 else L.TC_BF16
    d_q0 = torch.empty(Rp, D, device=DEV); d_kv0 = torch.empty(N * E, 2 * D, device=DEV)
    L.call("navc_cross_attention_bwd_tc", md, L.ptr(q), D, L.ptr(kv), 2 * D, P(seq_off), N, S, E, D, H, L.ptr(d_ctx), None, L.ptr(d_q0), D,
           L.ptr(d_kv0), 2 * D, L.stream())
    ld = Lw * 2 * D
    hi = torch.full((N * E, ld), 7.0, dtype=torch.bfloat16, device=DEV)
    lo = torch.full((N * E, ld), 7.0, dtype=torch.bfloat16, device=DEV) if x3 else None
    cs = torch.full((ld,), 0.5, device=DEV)
    d_q1 = torch.empty(Rp, D, device=DEV)
    off = 2 * D
    L.call("navc_cross_attention_bwd_tc_split", md, L.ptr(q), D, L.ptr(kv), 2 * D, P(seq_off), N, S, E, D, H, L.ptr(d_ctx), None,
           L.ptr(d_q1), D, hi[:, off:].data_ptr(), lo[:, off:].data_ptr() if x3 else None, ld, cs[off:].data_ptr(), L.stream())
    assert torch.equal(d_q0, d_q1)
    want_hi = d_kv0.to(torch.bfloat16)
    assert torch.equal(hi[:, off:off + 2 * D], want_hi)
    if x3:
        assert torch.equal(lo[:, off:off + 2 * D], (d_kv0 - want_hi.float()).to(torch.bfloat16))
    assert torch.all(hi[:, :off].float() == 7.0) and torch.all(hi[:, off + 2 * D:].float() == 7.0)
    close(cs[off:off + 2 * D] - 0.5, d_kv0.sum(0), 1e-5, "kv column sums")
    assert torch.all(cs[:off] == 0.5) and torch.all(cs[off + 2 * D:] == 0.5)


def test_packed_row_helpers():
    S, D, V = 7, 64, 301
    lens = [7, 2, 5, 1]
    N = len(lens)
    seq_off, rowmap64 = _pack(lens, S)
    rowmap = rowmap64.to(torch.int32).to(DEV)
    Rp, R = int(seq_off[-1]), N * S
    pad_rows = torch.tensor(sorted(set(range(R)) - set(rowmap64.tolist())), dtype=torch.int32, device=DEV)
    # gather / scatter of fp32 rows
    x = torch.randn(R, D, generator=g(66)).to(DEV)
    xp = torch.empty(Rp, D, device=DEV)
    L.call("navc_rows_f32", L.ptr(x), L.ptr(xp), D, L.ptr(rowmap), Rp, 0, L.stream())
    assert torch.equal(xp.cpu(), x.cpu()[rowmap64])
    back = torch.zeros(R, D, device=DEV)
    L.call("navc_rows_f32", L.ptr(xp), L.ptr(back), D, L.ptr(rowmap), Rp, 1, L.stream())
    want = torch.zeros(R, D)
    want[rowmap64] = x.cpu()[rowmap64]
    assert torch.equal(back.cpu(), want)
    # log-softmax of packed logits into padded rows; backward from padded g / logp into packed dlogits
    Vp = 320
    bias = torch.randn(V, generator=g(67))
    logits = torch.randn(Rp, Vp, generator=g(68)).to(DEV)
    const_lp = torch.log_softmax(bias, 0).view(1, V).to(DEV)
    out = const_lp.expand(R, V).contiguous()
    L.call("navc_log_softmax_rows", L.ptr(logits), Vp, L.ptr(out), V, L.ptr(rowmap), Rp, V, L.stream())
    want = const_lp.cpu().expand(R, V).clone()
    want[rowmap64] = torch.log_softmax(logits.cpu()[:, :V], -1)
    close(out, want, 1e-6, "log_softmax_rows")
    gr = torch.randn(R, V, generator=g(69))
    gr[pad_rows.cpu().long()[0]] = 0.0  # one all-zero PAD row (the PAD-ignoring loss case), the others not
    grd = gr.to(DEV)
    dlog = torch.empty(Rp, Vp, device=DEV)
    L.call("navc_log_softmax_bwd_rows", L.ptr(grd), L.ptr(out), L.ptr(rowmap), Rp, V, V, L.ptr(dlog), Vp, L.stream())
    full = gr - torch.exp(want) * gr.sum(-1, keepdim=True)  # d log_softmax, every padded row
    close(dlog[:, :V], full[rowmap64], 1e-5, "log_softmax_bwd_rows")
    assert dlog[:, V:].abs().max().item() == 0.0
    db = torch.zeros(V, device=DEV)
    L.call("navc_log_softmax_bwd_padrows", L.ptr(grd), V, L.ptr(pad_rows), pad_rows.numel(), L.ptr(const_lp), V, L.ptr(db),
           L.stream())
    close(db, full[pad_rows.cpu().long()].sum(0), 1e-5, "pad rows bias gradient")


GRAD_TOL_PACKED = 5e-4


@pytest.mark.parametrize("method,kw", [("NACF", {}), ("NAB", {}), ("ARB", {}), ("NACF", {"with_layernorm": True})],
                         ids=["nacf", "nab", "arb", "nacf_ln"])
def test_packed_training_matches_oracle_and_padded_path(method, kw):
    """dk == 64 (tcgen05 attention cores) -> the training path packs the real positions; loss and every gradient
    against the oracle, and against the padded path of the same build (navc_train_packed=0)."""
    kw = dict(kw, num_attention_heads=2)
    model, sd, bn_state, ref, ref_loss, res, loss = _train_case(method, "bf16x3", **kw)
    rows, padded = model.engine.last_train_rows
    assert rows < padded, "packing did not engage"
    model2, _, _, _, _, res2, loss2 = _train_case(method, "bf16x3", navc_train_packed=0, **kw)
    assert model2.engine.last_train_rows[0] == padded
    tol = GRAD_TOL_PACKED
    assert abs(loss.item() - ref_loss.item()) < tol * max(1.0, abs(ref_loss.item()))
    assert abs(loss.item() - loss2.item()) < 1e-5 * max(1.0, abs(loss2.item()))
    for a, b, c in zip(res["tgt_word_logprobs"], ref["tgt_word_logprobs"], res2["tgt_word_logprobs"]):
        assert (a.detach().cpu() - b.detach()).abs().max().item() < 5e-4
        # PAD rows included (log_softmax(bias)); real rows differ by bf16x3 rounding only (the softmax tiles differ)
        assert (a.detach() - c.detach()).abs().max().item() < 1e-4
    grads2 = dict(model2.named_parameters())
    checked = 0
    for name, p in model.named_parameters():
        rg = sd[name].grad
        if rg is None or rg.abs().max().item() == 0:
            assert p.grad is None or p.grad.abs().max().item() < 1e-6, name
            continue
        # atol: the key-bias gradients are analytically zero (softmax is invariant to a key bias); what is left is the
        # rounding of dS -- 2^-17 relative per element through the split-bf16 operands of the tcgen05 backward
        atol = 2e-7 + (1e-4 * sd[name.replace("key.bias", "key.weight")].grad.abs().max().item() if name.endswith("key.bias") else 0.0)
        close(p.grad, rg, tol, name, atol=atol)
        close(p.grad, grads2[name].grad, 1e-4, name + " (packed vs padded)", atol=atol)
        checked += 1
    assert checked > 20


def test_packed_vocab_bias_gradient_with_unmasked_loss():
    """A loss that does NOT ignore PAD positions: the skipped rows still reach the bias gradient."""
    out = {}
    for packed in (1, 0):
        opt = cases.small("NAB", hidden_dropout_prob=0.0, encoder_dropout=0.0, num_attention_heads=2, navc_train_packed=packed)
        torch.manual_seed(0)
        model = navc_b200.get_model(opt)
        model.load_state_dict(cases.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 11))
        model.to(DEV).train()
        model.set_precision("bf16x3")
        feats, category = cases.synth_inputs(opt, 5)
        toks = cases.synth_tokens(opt, 5, kind="nar")
        res = model(feats=[f.to(DEV) for f in feats], tgt_tokens=toks["tokens"].to(DEV), category=category.to(DEV))
        lp = res["tgt_word_logprobs"][0]
        w = torch.randn(lp.shape, generator=g(70)).to(DEV)
        (lp * w).sum().backward()
        out[packed] = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        out[packed, "rows"] = model.engine.last_train_rows
    assert out[1, "rows"][0] < out[0, "rows"][0]
    for k, v in out[0].items():
        # atol: the key-bias gradients are analytically zero (softmax is invariant to a key bias): what is left is rounding
        # noise of dS (2^-17 relative per element through the split-bf16 operands), measured against the key WEIGHT gradient
        atol = 2e-6 + (1e-4 * out[0][k.replace("key.bias", "key.weight")].abs().max().item() if k.endswith("key.bias") else 0.0)
        close(out[1][k], v, 1e-4, k, atol=atol)


# ---------------------------------------------------------------------------------------------------
# model level: loss + every parameter gradient vs the oracle (dropout off, BatchNorm batch statistics)
# ---------------------------------------------------------------------------------------------------
GRAD_TOL = {"fp32": 2e-4, "bf16x3": 5e-4}


def _train_case(method, precision, batch=6, **kw):
    opt = cases.small(method, hidden_dropout_prob=0.0, encoder_dropout=0.0, **kw)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd0 = cases.synth_state_dict(shapes, 11)
    model.load_state_dict(sd0)
    model.to(DEV).train()
    model.set_precision(precision)
    nar = O.is_nar(opt)
    feats, category = cases.synth_inputs(opt, batch)
    toks = cases.synth_tokens(opt, batch, kind="nar" if nar else "ar")
    dis = opt["decoder"] == "BertDecoderDisentangled"
    if nar:
        tgt = [toks["tokens_1"], toks["tokens"]] if dis else toks["tokens"]
        labels = [toks["labels_1"], toks["labels"]] if dis else toks["labels"]
        length_target = toks["length_target"]
    else:
        tgt = [toks["tokens"], toks["tokens"]] if dis else toks["tokens"]
        labels = [toks["labels"], toks["labels"]] if dis else toks["labels"]
        length_target = None
    # oracle (CPU, autograd)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd0.items()}
    bn_state = {}
    ref = O.model_forward(sd, opt, feats, tgt, category, training=True, bn_state=bn_state)
    ref_loss = O.criterion(opt, ref, labels, length_target)
    ref_loss.backward()
    # product path
    dev = lambda t: [x.to(DEV) for x in t] if isinstance(t, (list, tuple)) else t.to(DEV)
    res = model(feats=dev(feats), tgt_tokens=dev(tgt), category=category.to(DEV))
    loss = O.criterion(opt, res, dev(labels), None if length_target is None else length_target.to(DEV))
    loss.backward()
    return model, sd, bn_state, ref, ref_loss, res, loss


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("method,kw", [("NACF", {}), ("NAB", {}), ("ARB", {}), ("NACF", {"with_layernorm": True}),
                                       ("NAB", {"no_encoder_bn": True, "with_category": False})],
                         ids=["nacf", "nab", "arb", "nacf_ln", "nab_plain"])
def test_gradients_match_oracle(method, kw, precision):
    model, sd, bn_state, ref, ref_loss, res, loss = _train_case(method, precision, **kw)
    tol = GRAD_TOL[precision]
    assert abs(loss.item() - ref_loss.item()) < tol * max(1.0, abs(ref_loss.item()))
    for a, b in zip(res["tgt_word_logprobs"], ref["tgt_word_logprobs"]):
        assert (a.detach().cpu() - b.detach()).abs().max().item() < 5e-4
    if "pred_length" in ref:
        assert (res["pred_length"].detach().cpu() - ref["pred_length"].detach()).abs().max().item() < 5e-4
    checked = 0
    for name, p in model.named_parameters():
        rg = sd[name].grad
        if rg is None or rg.abs().max().item() == 0:
            assert p.grad is None or p.grad.abs().max().item() < 1e-6, name
            continue
        assert p.grad is not None, name
        # atol: gradients that are analytically zero (softmax is invariant to the key bias) are pure
        # rounding noise (~1e-8) on both sides
        close(p.grad, rg, tol, name, atol=zero_atol(name, lambda k: sd[k].grad))
        checked += 1
    assert checked > 20
    for k, v in bn_state.items():  # running statistics updated as nn.BatchNorm1d does
        close(dict(model.named_buffers())[k], v, 1e-5, k)


@pytest.mark.parametrize("method", ["NACF", "ARB"])
def test_training_step_releases_activations_without_gc(method):
    """No reference cycle through the autograd nodes: with the cyclic GC off, the memory held after a step
    (locals dropped, gradients cleared) does not grow from step to step."""
    import gc
    opt = cases.small(method, num_attention_heads=2)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(DEV).train()
    model.set_precision("bf16x3")
    nar = O.is_nar(opt)
    feats, category = cases.synth_inputs(opt, 6)
    toks = cases.synth_tokens(opt, 6, kind="nar" if nar else "ar")
    dis = opt["decoder"] == "BertDecoderDisentangled"
    tgt = [toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)] if (nar and dis) else toks["tokens"].to(DEV)
    feats = [f.to(DEV) for f in feats]
    category = category.to(DEV)

    def step():
        res = model(feats=feats, tgt_tokens=tgt, category=category)
        loss = sum(lp.sum() for lp in res["tgt_word_logprobs"])
        if "pred_length" in res:
            loss = loss + res["pred_length"].sum()
        loss.backward()
        for p in model.parameters():
            p.grad = None

    step()
    gc.collect()
    gc.disable()
    try:
        step()
        torch.cuda.synchronize()
        m1 = torch.cuda.memory_allocated()
        step()
        step()
        torch.cuda.synchronize()
        m2 = torch.cuda.memory_allocated()
    finally:
        gc.enable()
    assert m2 <= m1, "activations of finished steps are still referenced: %d -> %d bytes" % (m1, m2)


def test_direct_gradient_accumulation_matches_autograd():
    """With parallel.GradientAllReduce owning p.grad (views of one flat buffer) the weight / bias gradient kernels
    accumulate straight into it; same gradients as the plain autograd route, also over two micro-batches."""
    from navc_b200 import parallel
    grads = {}
    for direct in (False, True):
        opt = cases.small("NACF", hidden_dropout_prob=0.0, encoder_dropout=0.0, num_attention_heads=2)
        torch.manual_seed(0)
        model = navc_b200.get_model(opt)
        model.load_state_dict(cases.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 11))
        model.to(DEV).train()
        model.set_precision("bf16x3")
        dp = parallel.GradientAllReduce(model, broadcast=False) if direct else None
        assert (model.engine.grad_sink is not None) == direct
        if direct:   # query | key | value of every layer (weights and biases) and the K | V of all layers are ONE view each
            assert len(dp._group_views) >= 2 * (2 * opt["num_hidden_layers_decoder"] + 1)
            for ids, v in dp._group_views.items():
                members = [p for p in dp.params if id(p) in ids]
                assert v.numel() == sum(p.numel() for p in members)
                assert all(p.grad.data_ptr() >= v.data_ptr() and p.grad.data_ptr() < v.data_ptr() + 4 * v.numel() for p in members)
        if dp is not None:
            dp.zero_grad()
        for mb in range(2):
            feats, category = cases.synth_inputs(opt, 5, seed=100 + mb)
            toks = cases.synth_tokens(opt, 5, seed=200 + mb, kind="nar")
            res = model(feats=[f.to(DEV) for f in feats], tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)],
                        category=category.to(DEV))
            loss = O.criterion(opt, res, [toks["labels_1"].to(DEV), toks["labels"].to(DEV)], toks["length_target"].to(DEV))
            loss.backward()
        grads[direct] = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads[True]) == set(grads[False])
    for k, v in grads[False].items():
        close(grads[True][k], v, 2e-5, k, atol=zero_atol(k, lambda n: grads[False][n]))


def test_training_with_dropout_runs_and_is_seed_reproducible():
    opt = cases.small("NACF")  # reference dropout probabilities (0.5)
    feats, category = cases.synth_inputs(opt, 4)
    toks = cases.synth_tokens(opt, 4)
    outs = []
    for rep in range(3):
        torch.manual_seed(0)
        model = navc_b200.get_model(opt).to(DEV).train()
        torch.manual_seed(5 if rep < 2 else 6)
        res = model(feats=[f.to(DEV) for f in feats], tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)],
                    category=category.to(DEV))
        loss = O.criterion(opt, res, [toks["labels_1"].to(DEV), toks["labels"].to(DEV)], toks["length_target"].to(DEV))
        loss.backward()
        gn = torch.stack([p.grad.norm() for p in model.parameters() if p.grad is not None])
        assert torch.isfinite(loss) and torch.isfinite(gn).all()
        outs.append((loss.item(), gn.cpu()))
    # same torch seed -> same dropout masks: identical loss, gradients equal up to the summation
    # order of the atomic split-K / column-sum accumulations
    assert outs[0][0] == outs[1][0] and torch.allclose(outs[0][1], outs[1][1], rtol=1e-4)
    assert outs[0][0] != outs[2][0]


def test_train_step_updates_weights_and_eval_sees_them():
    """forward/backward/clip/Adam as misc/run.py:254-261, then the inference kernels pick up the new weights."""
    opt = cases.small("NACF", hidden_dropout_prob=0.0, encoder_dropout=0.0)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(DEV)
    optim = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-4)
    feats, category = cases.synth_inputs(opt, 6)
    toks = cases.synth_tokens(opt, 6)
    fd = [f.to(DEV) for f in feats]
    losses = []
    for it in range(4):
        model.train()
        optim.zero_grad()
        res = model(feats=fd, tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)], category=category.to(DEV))
        loss = O.criterion(opt, res, [toks["labels_1"].to(DEV), toks["labels"].to(DEV)], toks["length_target"].to(DEV))
        loss.backward()
        torch.nn.utils.clip_grad_value_(model.parameters(), 5)
        optim.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        res = model(feats=fd, tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)], category=category.to(DEV))
        ref = O.model_forward(sd, opt, feats, [toks["tokens_1"], toks["tokens"]], category)
    for a, b in zip(res["tgt_word_logprobs"], ref["tgt_word_logprobs"]):
        assert (a.cpu() - b).abs().max().item() < 5e-4


def test_fused_clip_adam_matches_torch():
    """navc_clip_adam (one launch over flat buffers) vs clip_grad_value_ + torch.optim.Adam(weight_decay)
    (misc/run.py:260-261, misc/optim.py:61-62) with the reference's per-step / per-epoch LR schedule."""
    from navc_b200 import optim as nopt
    opt = cases.small("NACF")
    opt.update(optim="adam", learning_rate=5e-3, minimum_learning_rate=5e-4, decay=0.9, weight_decay=5e-4, grad_clip=0.05)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(DEV)
    ref = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    tref = torch.optim.Adam(ref, lr=opt["learning_rate"], weight_decay=opt["weight_decay"])
    sched = nopt.get_optimizer(opt, model)
    lr = opt["learning_rate"]
    for it in range(5):
        sched.zero_grad()
        tref.zero_grad()
        gen = torch.Generator().manual_seed(50 + it)
        for p, r in zip(model.parameters(), ref):
            gr = (0.1 * torch.randn(p.shape, generator=gen)).to(DEV)
            p.grad.add_(gr)           # autograd-style in-place accumulation into the flat buffer views
            r.grad = gr.clone()
        sched.step()
        torch.nn.utils.clip_grad_value_(ref, opt["grad_clip"])
        for gp in tref.param_groups:
            gp["lr"] = lr
        tref.step()
        if it == 2:
            sched.epoch_update_learning_rate()
            lr = max(opt["minimum_learning_rate"], opt["decay"] * lr)
    for p, r in zip(model.parameters(), ref):
        close(p, r, 2e-5, "fused adam param")
    assert abs(sched.get_lr() - lr) < 1e-12
    assert model.engine._sig is None  # the engine repacks the updated weights at its next use


def test_train_steps_with_fused_optimizer_and_flat_gradients():
    from navc_b200 import optim as nopt, parallel
    opt = cases.small("NACF", hidden_dropout_prob=0.0, encoder_dropout=0.0)
    opt.update(optim="adam", learning_rate=5e-4, minimum_learning_rate=5e-5, decay=0.9, weight_decay=5e-4, grad_clip=5)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(DEV)
    dp = parallel.GradientAllReduce(model)
    sched = nopt.get_optimizer(opt, model, grads=dp)
    feats, category = cases.synth_inputs(opt, 6)
    toks = cases.synth_tokens(opt, 6)
    fd = [f.to(DEV) for f in feats]
    losses = []
    for it in range(4):
        model.train()
        sched.zero_grad()
        res = model(feats=fd, tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)], category=category.to(DEV))
        loss = O.criterion(opt, res, [toks["labels_1"].to(DEV), toks["labels"].to(DEV)], toks["length_target"].to(DEV))
        loss.backward()
        dp.allreduce()
        sched.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    model.eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        res = model(feats=fd, tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)], category=category.to(DEV))
        ref = O.model_forward(sd, opt, feats, [toks["tokens_1"], toks["tokens"]], category)
    for a, b in zip(res["tgt_word_logprobs"], ref["tgt_word_logprobs"]):
        assert (a.cpu() - b).abs().max().item() < 5e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_fused_cross_entropy_matches_oracle(precision):
    """opt['navc_fused_ce']: projection + log-softmax + masked NLL as one autograd node (no [B,S,V]
    log-prob tensor) behind navc_b200.misc.crit -- loss, every gradient and the meters vs the oracle."""
    from navc_b200.misc import crit as ncrit
    from navc_b200.training import FusedCEFn, LazyLogProbs
    opt = cases.small("NACF", hidden_dropout_prob=0.0, encoder_dropout=0.0, navc_fused_ce=True)
    opt.update(crit_key=[("tgt_word_logprobs", "tgt_word_labels"), ("pred_length", "tgt_length")],
               crit_name=["Cap Loss", "Length Loss"], crit_scale=[1.0, 1.0])
    torch.manual_seed(0)
    model = navc_b200.get_model(opt)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd0 = cases.synth_state_dict(shapes, 11)
    model.load_state_dict(sd0)
    model.to(DEV).train()
    model.set_precision(precision)
    B = 7
    feats, category = cases.synth_inputs(opt, B)
    toks = cases.synth_tokens(opt, B)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd0.items()}
    ref = O.model_forward(sd, opt, feats, [toks["tokens_1"], toks["tokens"]], category, training=True)
    ref_loss = O.criterion(opt, ref, [toks["labels_1"], toks["labels"]], toks["length_target"])
    ref_loss.backward()
    old_chunk, FusedCEFn.CHUNK = FusedCEFn.CHUNK, 40  # several backward chunks even at this small size
    try:
        res = model(feats=[f.to(DEV) for f in feats], tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)],
                    category=category.to(DEV))
        assert all(isinstance(x, LazyLogProbs) for x in res["tgt_word_logprobs"])
        res["tgt_word_labels"] = [toks["labels_1"].to(DEV), toks["labels"].to(DEV)]
        res["tgt_length"] = toks["length_target"].to(DEV)
        crit = ncrit.get_criterion(opt)
        crit.reset_loss_recorder()
        loss = crit.get_loss(res)
        loss.backward()
    finally:
        FusedCEFn.CHUNK = old_chunk
    tol = GRAD_TOL[precision]
    assert abs(loss.item() - ref_loss.item()) < tol * max(1.0, abs(ref_loss.item()))
    for name, p in model.named_parameters():
        rg = sd[name].grad
        if rg is None or rg.abs().max().item() == 0:
            continue
        close(p.grad, rg, tol, name, atol=zero_atol(name, lambda k: sd[k].grad))
    # meters: top-1 accuracy and perplexity from the fused statistics vs the oracle's log-probs
    names, info = crit.get_loss_info()
    lp = ref["tgt_word_logprobs"][1].detach()
    lab = toks["labels"]
    mask = lab.ne(0)
    ppl = math.exp(float(-(lp.gather(2, lab.unsqueeze(2)).squeeze(2) * mask).sum() / mask.sum()))
    acc1 = float((lp.argmax(-1)[mask] == lab[mask]).float().mean())
    got = dict(zip(names, info))
    assert abs(got["Perplexity"] - ppl) < 1e-3 * ppl and abs(got["Word Acc1"] - acc1) < 1e-6
    assert torch.allclose(res["tgt_word_logprobs"][1].materialize().detach().cpu(), lp, atol=5e-4)


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("rows,n_out,k_in,split", [(300, 512, 256, 1), (7680, 2048, 512, 4), (1000, 136, 520, 3), (70, 64, 64, 1)])
def test_wgrad_tc_mn_major(mode, tol, rows, n_out, k_in, split):
    """navc_wgrad_tc: dW += dY^T X straight from the row-major operands (MN-major tcgen05 operands)."""
    dy = torch.randn(rows, n_out, generator=g(60)) / math.sqrt(rows)
    x = torch.randn(rows, k_in, generator=g(61))
    base = torch.randn(n_out, k_in, generator=g(62))
    out = base.clone().to(DEV)
    sp = lambda t: (t.to(torch.bfloat16).to(DEV), (t - t.to(torch.bfloat16).float()).to(torch.bfloat16).to(DEV))
    dh, dl = sp(dy)
    xh, xl = sp(x)
    ep = L.Epilogue(None, None, None, 0, 0, L.ptr(out), None, None, k_in, 0, split, 1)
    L.call("navc_wgrad_tc", L.TC_BF16X3 if mode == "bf16x3" else L.TC_BF16, L.ptr(dh), L.ptr(dl), n_out, L.ptr(xh), L.ptr(xl), k_in,
           rows, n_out, k_in, ep, L.stream())
    ref = base.double() + dy.double().t() @ x.double()
    close(out, ref, tol, "wgrad")


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_tied_vocabulary_projection_forward_and_gradients(precision):
    """opt['tie_weights'] (models/seq2seq.py:30-33): tgt_word_prj shares the word-embedding Parameter (and gains a
    bias).  The engine must resolve both state_dict names, and the shared Parameter's gradient is the SUM of the
    embedding-lookup and the projection contributions."""
    opt = cases.small("NAB", hidden_dropout_prob=0.0, encoder_dropout=0.0, tie_weights=True, num_attention_heads=2)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt)
    emb_key = "decoder.embedding.word_embeddings.weight"
    assert model.tgt_word_prj.weight is dict(model.named_parameters())[emb_key]
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(cases.synth_state_dict(shapes, 13))
    with torch.no_grad():
        model.tgt_word_prj.bias.normal_(0.0, 0.05, generator=torch.Generator().manual_seed(5))
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.to(DEV)
    model.set_precision(precision)
    feats, category = cases.synth_inputs(opt, 6)
    toks = cases.synth_tokens(opt, 6, kind="nar")
    dev = lambda t: [x.to(DEV) for x in t] if isinstance(t, (list, tuple)) else t.to(DEV)
    # eval forward
    model.eval()
    with torch.no_grad():
        res = model(feats=dev(feats), tgt_tokens=toks["tokens"].to(DEV), category=category.to(DEV))
        ref = O.model_forward(sd0, opt, feats, toks["tokens"], category)
    assert (res["tgt_word_logprobs"][0].cpu() - ref["tgt_word_logprobs"][0]).abs().max().item() < 5e-4
    # training forward / backward
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd0.items()}
    sd["tgt_word_prj.weight"] = sd[emb_key]   # one leaf behind both names
    bn_state = {}
    ref = O.model_forward(sd, opt, feats, toks["tokens"], category, training=True, bn_state=bn_state)
    ref_loss = O.criterion(opt, ref, toks["labels"], toks["length_target"])
    ref_loss.backward()
    model.train()
    res = model(feats=dev(feats), tgt_tokens=toks["tokens"].to(DEV), category=category.to(DEV))
    loss = O.criterion(opt, res, toks["labels"].to(DEV), toks["length_target"].to(DEV))
    loss.backward()
    tol = GRAD_TOL[precision]
    assert abs(loss.item() - ref_loss.item()) < tol * max(1.0, abs(ref_loss.item()))
    checked = 0
    for name, p in model.named_parameters():
        rg = sd[name].grad
        if rg is None or rg.abs().max().item() == 0:
            continue
        assert p.grad is not None, name
        close(p.grad, rg, tol, name, atol=zero_atol(name, lambda k: sd[k].grad))
        checked += 1
    assert checked > 15 and model.tgt_word_prj.bias.grad is not None


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_layernorm_encoder_norm_matches_oracle(precision):
    """norm_type='ln' (models/joint_representation.py:20, 46-47): nn.LayerNorm per frame row instead of BatchNorm1d."""
    opt = cases.small("NACF", norm_type="ln", num_attention_heads=2, use_ct=True)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert "joint_representation_learner.ln0.weight" in shapes
    sd = cases.synth_state_dict(shapes, 17)
    model.load_state_dict(sd)
    model.to(DEV).eval()
    model.set_precision(precision)
    feats, category = cases.synth_inputs(opt, 6)
    with torch.no_grad():
        enc = model.encode(feats=[f.to(DEV) for f in feats])
        ref = O.encode(sd, opt, feats)
    tol = 1e-4 if precision == "fp32" else 5e-4
    for k in ("enc_output", "enc_hidden", "pred_length"):
        assert (enc[k].cpu() - ref[k]).abs().max().item() < tol, k
    hyp_o, det = O.translate(sd, opt, feats, category, return_details=True)
    with torch.no_grad():
        hyp, _ = navc_b200.Translator(model, opt, device=DEV).translate_batch(enc, category.to(DEV), None, {})
    for b in (hyp.cpu() != hyp_o).any(1).nonzero().flatten().tolist():
        assert det["video_margin"][b].item() <= 1e-4, (b, det["video_margin"][b].item())


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_weight_refresh_equals_full_repack(precision):
    """After an in-place parameter update (an optimizer step, load_state_dict) the engine rewrites its packed operands
    with ONE navc_refresh_pack launch instead of re-running the packing; the result must be bit-identical to a full
    repack, for concatenated operands (q|k|v, all-layer K|V, highway w1|w2), single ones and their bf16 hi/lo copies."""
    opt = cases.small("NACF", num_attention_heads=2)
    torch.manual_seed(0)
    model = navc_b200.get_model(opt)
    model.load_state_dict(cases.synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 11))
    model.to(DEV).eval()
    model.set_precision(precision)
    feats, category = cases.synth_inputs(opt, 5)
    toks = cases.synth_tokens(opt, 5, kind="nar")
    args = dict(feats=[f.to(DEV) for f in feats], tgt_tokens=[toks["tokens_1"].to(DEV), toks["tokens"].to(DEV)], category=category.to(DEV))
    with torch.no_grad():
        out0 = model(**args)["tgt_word_logprobs"][1].clone()
        eng = model.engine
        pid = eng.pack_id
        assert eng._refresh_table is not None and eng._refresh_table.shape[1] == 5
        g = torch.Generator().manual_seed(3)
        for p in model.parameters():
            p.add_((torch.randn(p.shape, generator=g) * 0.01).to(DEV))
        launches = L.launches
        out1 = model(**args)["tgt_word_logprobs"][1].clone()       # refresh path
        assert eng.pack_id == pid + 1
        eng._refresh_table = None                                    # force the full repack of the same weights
        eng.invalidate()
        out2 = model(**args)["tgt_word_logprobs"][1].clone()
        assert eng.pack_id == pid + 2
    assert (out1 - out0).abs().max().item() > 1e-4                   # the update was seen
    assert torch.equal(out1, out2)
