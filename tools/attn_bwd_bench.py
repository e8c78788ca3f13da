#!/usr/bin/env python
"""Device time of the attention backward kernels at the training shapes (B=256, S=30, E=120, D=512, H=8),
padded and packed.  python tools/attn_bwd_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import navc_b200  # noqa: E402,F401
from navc_b200 import _lib as L  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    L.ensure_init(dev)
    N, S, E, D, H = 256, 30, 120, 512, 8
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(4, 29, (N,), generator=g)
    seq_off = torch.zeros(N + 1, dtype=torch.int32)
    seq_off[1:] = torch.cumsum(lens, 0)
    Rp = int(seq_off[-1])
    toks = torch.zeros(N, S, dtype=torch.int64)
    for n in range(N):
        toks[n, :lens[n]] = 7
    toks, seq_off = toks.to(dev), seq_off.to(dev)
    R = N * S
    qkv = torch.randn(R, 3 * D, device=dev)
    d_ctx = torch.randn(R, D, device=dev)
    d_qkv = torch.empty(R, 3 * D, device=dev)
    q = torch.randn(R, D, device=dev)
    kv = torch.randn(N * E, 12 * D, device=dev)   # the per-layer K|V slice of the all-layer projection (ld = 12 D)
    d_q = torch.empty(R, D, device=dev)
    d_kv = torch.empty(N * E, 12 * D, device=dev)
    st = L.stream()
    res = {}
    res["self padded"] = timeit(lambda: L.call("navc_self_attention_bwd", L.ptr(qkv), 3 * D, L.ptr(toks), N, S, D, H, 0, 0,
                                               L.ptr(d_ctx), L.ptr(d_qkv), st))
    res["self packed"] = timeit(lambda: L.call("navc_self_attention_bwd_packed", L.ptr(qkv), 3 * D, L.ptr(toks), L.ptr(seq_off), N, S,
                                               D, H, 0, 0, L.ptr(d_ctx), L.ptr(d_qkv), st))
    res["cross padded"] = timeit(lambda: L.call("navc_cross_attention_bwd", L.ptr(q), D, L.ptr(kv), 12 * D, N, S, E, D, H, 1,
                                                L.ptr(d_ctx), L.ptr(d_q), D, L.ptr(d_kv), 12 * D, st))
    res["cross packed"] = timeit(lambda: L.call("navc_cross_attention_bwd_packed", L.ptr(q), D, L.ptr(kv), 12 * D, L.ptr(seq_off), N, S,
                                                E, D, H, L.ptr(d_ctx), L.ptr(d_q), D, L.ptr(d_kv), 12 * D, st))
    print("rows packed %d of %d" % (Rp, R))
    for k, v in res.items():
        print("%-14s %8.1f us" % (k, v))


if __name__ == "__main__":
    main()
