#!/usr/bin/env python
"""Isolated timing of navc_linear_tf32 at the decoder-layer shapes (fp32 in, bf16 hi/lo or fp32 out), graph-free CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L
dev = torch.device("cuda", 0); L.ensure_init(dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10553
for name, N, K in (("qkv", 1536, 512), ("cq", 512, 512), ("f1", 2048, 512), ("f2", 512, 2048)):
    xs = [torch.randn(M, K, device=dev) for _ in range(3)]
    w = torch.randn(N, K, device=dev) * 0.05; b = torch.randn(N, device=dev)
    hi = torch.empty(M, N, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi); out = torch.empty(M, N, device=dev)
    for label, ep in (("hi/lo out", L.Epilogue(L.ptr(b), None, None, 0, 0, None, L.ptr(hi), L.ptr(lo), N, 0, 1, 0, None, None, None, 0, 0)),
                      ("fp32 out", L.Epilogue(L.ptr(b), None, None, 0, 0, L.ptr(out), None, None, N, 0, 1, 0, None, None, None, 0, 0))):
        run = lambda i: L.call("navc_linear_tf32", L.ptr(xs[i % 3]), K, L.ptr(w), K, M, N, K, ep, L.stream())
        for i in range(3): run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): run(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 50
        print("tf32 %-4s M=%d N=%4d K=%4d %-9s %7.1f us  %6.1f TFLOP/s" % (name, M, N, K, label, us, 2.0 * M * N * K / us / 1e6))
