"""Second-generation attention cores (csrc/attention2_tc.cu: persistent, pipelined, 96-row windows of the packed row
space) through the C ABI against a plain torch fp32 reference of the same op (models/bert.py:139-179): per (row,
head) softmax(mask_fill(Q K^T / sqrt(dk), -1e7)) V.  Tolerances: bf16x3 3e-5 of the output magnitude, bf16 3e-2."""
import math

import pytest
import torch

import navc_b200
from navc_b200 import _lib as L
from oracle import navc_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def g(seed):
    return torch.Generator().manual_seed(seed)


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def packing(N, S, seed, lo=1):
    lens = torch.randint(lo, S + 1, (N,), generator=g(seed)).int()
    off = torch.zeros(N + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(lens, 0)
    return lens, off


def tiles(off_d, N, S):
    L.ensure_init(DEV)
    w = L._lib.navc_attention_window()
    n_tiles = (N * S + w - 1) // w
    tile_seq = torch.full((n_tiles + 1,), -7, dtype=torch.int32, device=DEV)
    L.call("navc_pack_tiles", L.ptr(off_d), N, L.ptr(tile_seq), n_tiles, L.stream())
    return tile_seq, n_tiles


def test_pack_tiles_windows():
    N, S = 300, 32
    lens, off = packing(N, S, 5)
    off_d = off.to(DEV)
    tile_seq, n_tiles = tiles(off_d, N, S)
    ts = tile_seq.cpu().tolist()
    w = L._lib.navc_attention_window()
    assert ts[0] == 0 and ts[-1] == N and all(a <= b for a, b in zip(ts, ts[1:]))
    for t in range(n_tiles):
        for n in range(ts[t], ts[t + 1]):
            assert t * w <= int(off[n]) < (t + 1) * w            # every sequence lives in the window its first row is in
        if ts[t + 1] > ts[t]:
            assert int(off[ts[t + 1]]) - int(off[ts[t]]) <= 128  # a tile's rows fit one UMMA M = 128


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("kind", ["NARFormer", "ARFormer", "SelfMask"])
@pytest.mark.parametrize("N,S,lo", [(7, 11, 1), (45, 27, 4), (768, 28, 4), (130, 32, 30), (3000, 29, 1)])
def test_self_attention_tiles(mode, tol, kind, N, S, lo):
    D, H = 512, 8
    lens, off = packing(N, S, 46, lo)
    R = int(off[-1])
    gen = g(47)
    qkv = torch.randn(N * S, 3 * D, generator=gen)          # packed rows first, NaN beyond (rows past the count are never written)
    qkv[R:] = float("nan")
    toks = torch.zeros(N, S, dtype=torch.int64)
    for n in range(N):
        toks[n, :lens[n]] = torch.randint(1, 50, (int(lens[n]),), generator=gen)
    if lens[1] > 2:
        toks[1, 1] = 0                                       # an interior (predicted) <pad>
    hi, lo_ = split(qkv)
    x3 = mode == "bf16x3"
    chi = torch.zeros(N * S, D, dtype=torch.bfloat16, device=DEV)
    clo = torch.zeros(N * S, D, dtype=torch.bfloat16, device=DEV) if x3 else None
    hi_d, lo_d, toks_d, off_d = hi.to(DEV), lo_.to(DEV), toks.to(DEV), off.to(DEV)
    tile_seq, n_tiles = tiles(off_d, N, S)
    for rep in range(2):
        L.call("navc_self_attention_tc_tiles", L.TC_BF16X3 if x3 else L.TC_BF16, L.ptr(hi_d), L.ptr(lo_d) if x3 else None,
               3 * D, L.ptr(toks_d), L.ptr(off_d), L.ptr(tile_seq), n_tiles, N * S, N, S, D, H, L.MASK_KIND[kind], 0,
               L.ptr(chi), L.ptr(clo), L.stream())
    torch.cuda.synchronize()
    ctx = (chi.float() + (clo.float() if x3 else 0)).cpu()
    qe = (hi.float() + (lo_.float() if x3 else 0))           # the operand values the kernel sees
    dk = D // H
    worst = 0.0
    step = max(1, N // 60)
    for n in list(range(0, N, step)) + [N - 1]:
        ln, r0 = int(lens[n]), int(off[n])
        blk = qe[r0:r0 + ln].double()
        q, k, v = [t.view(ln, H, dk).permute(1, 0, 2) for t in blk.split(D, dim=1)]
        mask = O.self_attention_mask(toks[n:n + 1, :ln], kind, 0)[0]
        sc = ((q @ k.transpose(-1, -2)) / math.sqrt(dk)).masked_fill(mask.unsqueeze(0), O.MASK_FILL)
        ref = (torch.softmax(sc, -1) @ v).permute(1, 0, 2).reshape(ln, D)
        worst = max(worst, (ctx[r0:r0 + ln].double() - ref).abs().max().item() / max(1.0, ref.abs().max().item()))
    assert worst < tol, worst
    assert ctx[R:].abs().max().item() == 0                   # rows beyond the packed count are never written


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("bf16", 3e-2)])
@pytest.mark.parametrize("B,group,S,E,lo", [(3, 1, 9, 16, 1), (5, 6, 28, 120, 4), (128, 6, 28, 120, 4), (2, 10, 30, 128, 20), (40, 3, 12, 37, 1)])
def test_cross_attention_tiles(mode, tol, B, group, S, E, lo):
    D, H = 512, 8
    N = B * group
    lens, off = packing(N, S, 48, lo)
    R = int(off[-1])
    gen = g(49)
    q = torch.randn(N * S, D, generator=gen)
    q[R:] = float("nan")                                     # rows past the packed count are never written
    kv = torch.randn(B * E, 2 * D, generator=gen)
    qh, ql = split(q)
    kh, kl_ = split(kv)
    x3 = mode == "bf16x3"
    chi = torch.zeros(N * S, D, dtype=torch.bfloat16, device=DEV)
    clo = torch.zeros(N * S, D, dtype=torch.bfloat16, device=DEV) if x3 else None
    qh_d, ql_d, kh_d, kl_d, off_d = qh.to(DEV), ql.to(DEV), kh.to(DEV), kl_.to(DEV), off.to(DEV)
    L.ensure_init(DEV)
    for rep in range(2):
        L.call("navc_cross_attention_tc_tiles", L.TC_BF16X3 if x3 else L.TC_BF16, L.ptr(qh_d), L.ptr(ql_d) if x3 else None, D,
               L.ptr(kh_d), L.ptr(kl_d) if x3 else None, 2 * D, L.ptr(off_d), N * S, N, S, E, D, H, group, L.ptr(chi), L.ptr(clo),
               L.stream())
    torch.cuda.synchronize()
    ctx = (chi.float() + (clo.float() if x3 else 0)).cpu()
    qe = (qh.float() + (ql.float() if x3 else 0)).double()
    kve = (kh.float() + (kl_.float() if x3 else 0)).double()
    dk = D // H
    worst = 0.0
    step = max(1, N // 60)
    for n in list(range(0, N, step)) + [N - 1]:
        ln, r0, b = int(lens[n]), int(off[n]), n // group
        qq = qe[r0:r0 + ln].view(ln, H, dk).permute(1, 0, 2)
        kk = kve[b * E:(b + 1) * E, :D].view(E, H, dk).permute(1, 0, 2)
        vv = kve[b * E:(b + 1) * E, D:].view(E, H, dk).permute(1, 0, 2)
        ref = (torch.softmax(qq @ kk.transpose(-1, -2) / math.sqrt(dk), -1) @ vv).permute(1, 0, 2).reshape(ln, D)
        worst = max(worst, (ctx[r0:r0 + ln].double() - ref).abs().max().item() / max(1.0, ref.abs().max().item()))
    assert worst < tol, worst
    assert ctx[R:].abs().max().item() == 0
