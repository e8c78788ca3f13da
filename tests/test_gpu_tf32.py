"""kind::tf32 GEMM (csrc/gemm_tc.cu, navc_linear_tf32) through the C ABI against a float64 torch reference, and the engine's
'tf32' precision mode end to end.  Tolerance: TF32 keeps 10 explicit mantissa bits of each operand (2^-11 relative per
element after the TMA's round-to-nearest conversion -- CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 --, random signs over the K terms):
1e-3 of the output magnitude (measured ~3e-4; plain FLOAT32 maps let the tensor core truncate: 8x the error, biased)."""
import math

import pytest
import torch

import cases
import navc_b200
from navc_b200 import _lib as L
from oracle import navc_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("M,N,K", [(200, 512, 512), (4195, 2048, 512), (333, 1536, 512), (77, 512, 2048), (130, 30, 96), (1, 64, 36),
                                   (10553, 512, 512)])
@pytest.mark.parametrize("epi", ["plain", "full"])
def test_linear_tf32_matches_fp64(M, N, K, epi):
    L.ensure_init(DEV)
    x = torch.randn(M, K, generator=g(1))
    w = torch.randn(N, K, generator=g(2)) / math.sqrt(K)
    b = torch.randn(N, generator=g(3))
    res = torch.randn(M, N, generator=g(4))
    toks = torch.randint(0, 4, (M,), generator=g(5))
    y = x.double() @ w.double().t()
    full = epi == "full"
    if full:
        y = O.activation("gelu_new")((y + b.double()).float()).double()
        y = (y + res.double()) * toks.ne(0).double().unsqueeze(1)
    xd, wd, bd, rd, td = x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV), toks.to(DEV)
    ldo = (N + 7) // 8 * 8
    out = torch.full((M, ldo), float("nan"), device=DEV)
    hi = torch.zeros((M, ldo), dtype=torch.bfloat16, device=DEV)
    lo = torch.zeros((M, ldo), dtype=torch.bfloat16, device=DEV)
    ep = L.Epilogue(L.ptr(bd) if full else None, L.ptr(rd) if full else None, L.ptr(td) if full else None, 1 if full else 0,
                    N if full else 0, L.ptr(out), L.ptr(hi), L.ptr(lo), ldo, 0, 1, 0, None, None, None, 0, 0)
    L.call("navc_linear_tf32", L.ptr(xd), K, L.ptr(wd), K, M, N, K, ep, L.stream())
    got = out[:, :N].cpu().double()
    scale = max(y.abs().max().item(), 1e-6)
    assert (got - y).abs().max().item() < 1e-3 * scale
    pair = (hi.float() + lo.float())[:, :N].cpu().double()
    assert (pair - got).abs().max().item() < 1e-4 * scale   # bf16 hi + lo carries the fp32 result


def test_linear_tf32_device_side_row_count():
    L.ensure_init(DEV)
    M, N, K, live = 700, 512, 512, 389
    x = torch.randn(M, K, generator=g(7)).to(DEV)
    w = (torch.randn(N, K, generator=g(8)) / math.sqrt(K)).to(DEV)
    out = torch.full((M, N), 7.0, device=DEV)
    cnt = torch.tensor([live], dtype=torch.int32, device=DEV)
    ep = L.Epilogue(None, None, None, 0, 0, L.ptr(out), None, None, N, 0, 1, 0, None, None, cnt.data_ptr(), 0, 0)
    L.call("navc_linear_tf32", L.ptr(x), K, L.ptr(w), K, M, N, K, ep, L.stream())
    want = x[:live].double() @ w.double().t()
    assert (out[:live].double() - want).abs().max().item() < 1e-3 * want.abs().max().item()
    assert torch.all(out[(live + 127) // 128 * 128:] == 7.0)   # whole tiles beyond the count are never touched


@pytest.mark.parametrize("kw", [dict(paradigm="mp", use_ct=True), dict(paradigm="ef", use_ct=True, q=2)])
def test_tf32_mode_decodes_like_the_oracle_where_margins_allow(kw):
    """The engine's 'tf32' mode (QKV / query / FFN / K|V / encoder projections on kind::tf32, the rest split-bf16) at the
    headline head size: ids equal the oracle's for every video whose smallest decision margin exceeds 5e-3 log units,
    through packed rows and the replayed CUDA graph; log-probabilities within 2e-2."""
    opt = cases.wide("NACF", **kw)
    model = navc_b200.get_model(opt)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = cases.synth_state_dict(shapes, 11, 0.5)
    model.load_state_dict(sd)
    model.to(DEV).eval()
    model.set_precision("tf32")
    assert model.engine.tf32 and model.engine.split
    feats, category = cases.synth_inputs(opt, 9)
    hyp_o, det = O.translate(sd, opt, feats, category, return_details=True)
    tr = navc_b200.Translator(model, opt, device=DEV)
    with torch.no_grad():
        for _ in range(3):
            enc = model.encode(feats=[f.to(DEV) for f in feats])
            hyp, _ = tr.translate_batch(enc, category.to(DEV), None, {})
    st = navc_b200.generate.last_stats
    assert st["packed"]
    hyp = hyp.cpu()
    eq = (hyp == hyp_o).all(1)
    above = det["video_margin"] > 5e-3
    assert not (~eq & above).any(), (det["video_margin"].tolist(), eq.tolist())
    assert eq.float().mean().item() >= 0.5
