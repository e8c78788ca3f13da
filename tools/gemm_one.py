#!/usr/bin/env python
"""Launch a few navc_linear_tc shapes (pair epilogue) for ncu captures: python tools/gemm_one.py bf16|bf16x3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L
dev = torch.device("cuda", 0)
L.ensure_init(dev)
mode = L.TC_BF16X3 if (len(sys.argv) > 1 and sys.argv[1] == "bf16x3") else L.TC_BF16
m = 21504
for name, N, K, act, res in (("f1", 2048, 512, 1, False), ("so", 512, 512, 0, True), ("qkv", 1536, 512, 0, False)):
    xh = torch.randn(m, K, device=dev).to(torch.bfloat16); xl = xh * 0.01
    wh = torch.randn(N, K, device=dev).to(torch.bfloat16); wl = wh * 0.01
    b = torch.randn(N, device=dev)
    rh = torch.randn(m, N, device=dev).to(torch.bfloat16) if res else None
    rl = (rh * 0.01) if res else None
    toks = torch.ones(m, dtype=torch.int64, device=dev)
    oh = torch.empty(m, N, dtype=torch.bfloat16, device=dev); ol = torch.empty_like(oh)
    ep = L.Epilogue(L.ptr(b), None, L.ptr(toks) if res else None, act, N if res else 0, None, L.ptr(oh), L.ptr(ol), N, 0, 1, 0, L.ptr(rh), L.ptr(rl))
    for i in range(3):
        L.call("navc_linear_tc", mode, L.ptr(xh), L.ptr(xl), K, L.ptr(wh), L.ptr(wl), K, m, N, K, ep, L.stream())
    torch.cuda.synchronize()
