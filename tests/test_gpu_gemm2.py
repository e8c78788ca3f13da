"""Second-generation tcgen05 GEMM (csrc/gemm2_tc.cu: 2-CTA clusters multicasting the weight tile, tail split along N,
pair epilogue) through the C ABI (navc_linear_tc with bf16-only outputs) against a float64 torch reference of the same op.

Tolerances: bf16x3 (split-bf16 operands, 3 products) 3e-5 of the output magnitude; bf16 3e-2 (operand rounding)."""
import math
import os

import pytest
import torch

import navc_b200
from navc_b200 import _lib as L
from oracle import navc_oracle as O

os.environ["NAVC_GEMM2"] = "force"   # read once by the library: gemm2 also in plain bf16 mode (the product uses it for bf16x3)
pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def g(seed):
    return torch.Generator().manual_seed(seed)


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


@pytest.fixture(autouse=True)
def _init():
    L.ensure_init(DEV)
    yield
    os.environ.pop("NAVC_GEMM2_CLUSTER", None)


def run(mode, M, N, K, cluster, force=0, cnt=None, epilogue=True, seed=60, act="none"):
    """epilogue=True: bias + `act` + bf16-pair residual + row mask (act='none' is what the decoder layer's residual
    GEMMs use and what gemm2 serves: the residual is preloaded into the accumulators; an activation together with a
    residual goes to the first-generation kernel)."""
    x = torch.randn(M, K, generator=g(seed))
    w = torch.randn(N, K, generator=g(seed + 1)) / math.sqrt(K)
    b = torch.randn(N, generator=g(seed + 2))
    res = torch.randn(M, N, generator=g(seed + 3))
    toks = torch.randint(0, 4, (M,), generator=g(seed + 4))
    x3 = mode == "bf16x3"
    xh, xl = [t.to(DEV) for t in split(x)]
    wh, wl = [t.to(DEV) for t in split(w)]
    rh, rl = [t.to(DEV) for t in split(res)]
    bd, td = b.to(DEV), toks.to(DEV)
    xe = xh.float().cpu().double() + (xl.float().cpu().double() if x3 else 0)
    we = wh.float().cpu().double() + (wl.float().cpu().double() if x3 else 0)
    y = xe @ we.t()
    if epilogue:
        y = y + b.double()
        if act != "none":
            y = O.activation(act)(y.float()).double()
        y = (y + (rh.float() + (rl.float() if x3 else 0)).cpu().double()) * toks.ne(0).double().unsqueeze(1)
    if cluster:
        os.environ["NAVC_GEMM2_CLUSTER"] = str(cluster)
    ohi = torch.full((M, N), 9.0, dtype=torch.bfloat16, device=DEV)
    olo = torch.full((M, N), 9.0, dtype=torch.bfloat16, device=DEV) if x3 else None
    m_dev = torch.tensor([cnt], dtype=torch.int32, device=DEV) if cnt is not None else None
    if epilogue:
        ep = L.Epilogue(L.ptr(bd), None, L.ptr(td), L.ACT[act], N, None, L.ptr(ohi), L.ptr(olo), N, force, 1, 0,
                        L.ptr(rh), L.ptr(rl) if x3 else None, L.ptr(m_dev), cnt or 0, 0)
    else:
        ep = L.Epilogue(None, None, None, 0, 0, None, L.ptr(ohi), L.ptr(olo), N, force, 1, 0, None, None, L.ptr(m_dev), cnt or 0, 0)
    for rep in range(2):
        L.call("navc_linear_tc", L.TC_BF16X3 if x3 else L.TC_BF16, L.ptr(xh), L.ptr(xl) if x3 else None, K, L.ptr(wh),
               L.ptr(wl) if x3 else None, K, M, N, K, ep, L.stream())
    torch.cuda.synchronize()
    got = (ohi.float() + (olo.float() if x3 else 0)).cpu().double()
    rows = M if cnt is None else cnt
    tol = (3e-5 if x3 else 1.2e-2) * max(1.0, y.abs().max().item())
    err = (got[:rows] - y[:rows]).abs().max().item()
    assert err < tol, (mode, M, N, K, cluster, force, cnt, err)
    if cnt is not None:   # rows of m-blocks beyond the device-side count keep their old contents
        tile_end = (cnt + 127) // 128 * 128
        assert (ohi[tile_end:].float() == 9.0).all()


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("M,N,K", [(10478, 512, 512), (10478, 1536, 512), (10478, 2048, 512), (10478, 512, 2048),
                                   (3000, 520, 200), (129, 1024, 64), (7777, 264, 1032), (40000, 512, 128)])
def test_gemm2_matches_reference(mode, cluster, M, N, K):
    run(mode, M, N, K, cluster)
    run(mode, M, N, K, cluster, act="gelu_new", seed=70)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("act", ["gelu_new", "relu"])
def test_gemm2_activation_without_residual(mode, act):
    """bias + activation, no residual (the FFN up-projection): gemm2's plain epilogue."""
    M, N, K = 10478, 2048, 512
    x = torch.randn(M, K, generator=g(80))
    w = torch.randn(N, K, generator=g(81)) / math.sqrt(K)
    b = torch.randn(N, generator=g(82))
    x3 = mode == "bf16x3"
    xh, xl = [t.to(DEV) for t in split(x)]
    wh, wl = [t.to(DEV) for t in split(w)]
    xe = xh.float().cpu().double() + (xl.float().cpu().double() if x3 else 0)
    we = wh.float().cpu().double() + (wl.float().cpu().double() if x3 else 0)
    y = O.activation(act)((xe @ we.t() + b.double()).float()).double()
    ohi = torch.empty((M, N), dtype=torch.bfloat16, device=DEV)
    olo = torch.empty((M, N), dtype=torch.bfloat16, device=DEV) if x3 else None
    bd = b.to(DEV)
    ep = L.Epilogue(L.ptr(bd), None, None, L.ACT[act], 0, None, L.ptr(ohi), L.ptr(olo), N, 0, 1, 0, None, None, None, 0, 0)
    L.call("navc_linear_tc", L.TC_BF16X3 if x3 else L.TC_BF16, L.ptr(xh), L.ptr(xl) if x3 else None, K, L.ptr(wh),
           L.ptr(wl) if x3 else None, K, M, N, K, ep, L.stream())
    got = (ohi.float() + (olo.float() if x3 else 0)).cpu().double()
    assert (got - y).abs().max().item() < (3e-5 if x3 else 1.2e-2) * max(1.0, y.abs().max().item())


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("force", [7, 128, 256])   # 7: tail split off; 128 / 256: forced tile width
@pytest.mark.parametrize("cluster", [1, 2])
def test_gemm2_tile_shapes_and_tail_split(mode, force, cluster):
    run(mode, 10478, 512, 512, cluster, force=force)
    run(mode, 5200, 768, 256, cluster, force=force, epilogue=False)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("cnt", [10478, 1, 129, 21504])
def test_gemm2_device_row_count(mode, cluster, cnt):
    """Packed rows: launch sized for 21504 rows, only *m_dev of them computed (odd m-block counts leave the second
    CTA of the last cluster without a tile)."""
    run(mode, 21504, 512, 512, cluster, cnt=cnt)
    run(mode, 21504, 2048, 512, cluster, cnt=cnt, epilogue=False)


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,live", [(21504, 10553), (21504, 3285), (700, 700), (130, 77), (21504, 21504)])
def test_chained_linears_equal_two_launches(mode, M, live):
    """navc_linear_chain_tc (out-projection + residual, then the next projection, in one persistent launch with per-row-block
    completion counters) against the same two layers as separate launches: equal outputs, over repeated launches with fresh
    inputs (the second problem's tiles must never read a row block of y0 before it is complete)."""
    D = 512
    x3 = mode == "bf16x3"
    md = L.TC_BF16X3 if x3 else L.TC_BF16
    cnt = torch.tensor([live], dtype=torch.int32, device=DEV)
    w0 = torch.randn(D, D, generator=g(1)) / math.sqrt(D); w1 = torch.randn(D, D, generator=g(2)) / math.sqrt(D)
    b0 = torch.randn(D, generator=g(3)).to(DEV); b1 = torch.randn(D, generator=g(4)).to(DEV)
    w0h, w0l = [t.to(DEV) for t in split(w0)]; w1h, w1l = [t.to(DEV) for t in split(w1)]
    for rep in range(6):
        x = torch.randn(M, D, generator=g(10 + rep)); res = torch.randn(M, D, generator=g(40 + rep))
        xh, xl = [t.to(DEV) for t in split(x)]; rh, rl = [t.to(DEV) for t in split(res)]
        outs = []
        for chained in (False, True):
            a_h = torch.full((M, D), float("nan"), dtype=torch.bfloat16, device=DEV); a_l = torch.full_like(a_h, float("nan"))
            q_h = torch.full((M, D), float("nan"), dtype=torch.bfloat16, device=DEV); q_l = torch.full_like(q_h, float("nan"))
            e0 = L.Epilogue(L.ptr(b0), None, None, 0, D, None, L.ptr(a_h), L.ptr(a_l) if x3 else None, D, 0, 1, 0, L.ptr(rh), L.ptr(rl) if x3 else None,
                            cnt.data_ptr(), live, 0)
            e1 = L.Epilogue(L.ptr(b1), None, None, 0, 0, None, L.ptr(q_h), L.ptr(q_l) if x3 else None, D, 0, 1, 0, None, None, cnt.data_ptr(), live, 0)
            if chained:
                L.call("navc_linear_chain_tc", md, L.ptr(xh), L.ptr(xl) if x3 else None, D, L.ptr(w0h), L.ptr(w0l) if x3 else None, D, e0,
                       L.ptr(w1h), L.ptr(w1l) if x3 else None, D, e1, M, D, D, L.stream())
            else:
                L.call("navc_linear_tc", md, L.ptr(xh), L.ptr(xl) if x3 else None, D, L.ptr(w0h), L.ptr(w0l) if x3 else None, D, M, D, D, e0, L.stream())
                L.call("navc_linear_tc", md, L.ptr(a_h), L.ptr(a_l) if x3 else None, D, L.ptr(w1h), L.ptr(w1l) if x3 else None, D, M, D, D, e1, L.stream())
            torch.cuda.synchronize()
            outs.append([t[:live].float().cpu() for t in ((a_h, a_l, q_h, q_l) if x3 else (a_h, q_h))])
        for u, v in zip(*outs):
            assert not torch.isnan(v).any()
            assert torch.equal(u, v)
