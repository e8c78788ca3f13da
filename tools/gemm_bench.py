#!/usr/bin/env python
"""Per-shape timing of navc_linear_tc (CUDA events, L2-cold via rotating operands)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L

dev = torch.device("cuda", 0)
L.ensure_init(dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 21504
DBG = int(os.environ.get("NAVC_DBG", "0"))
shapes = [("qkv", 1536, 512, 0, False), ("so", 512, 512, 0, True), ("f1", 2048, 512, 1, False), ("f2", 512, 2048, 0, True),
          ("enc0", 512, 2048, 0, False)]
for mode, mname in ((L.TC_BF16X3, "bf16x3"), (L.TC_BF16, "bf16")):
    for name, N, K, act, res in shapes:
        m = 15360 if name == "enc0" else M
        nb = 3
        xs = [torch.randn(m, K, device=dev).to(torch.bfloat16) for _ in range(nb)]
        xl = [torch.randn(m, K, device=dev).to(torch.bfloat16) * 0.01 for _ in range(nb)]
        wh = torch.randn(N, K, device=dev).to(torch.bfloat16); wl = wh * 0.01
        b = torch.randn(N, device=dev)
        r = torch.randn(m, N, device=dev) if res else None
        toks = torch.ones(m, dtype=torch.int64, device=dev)
        of = torch.empty(m, N, device=dev) if res else None
        oh = torch.empty(m, N, dtype=torch.bfloat16, device=dev)
        ol = torch.empty(m, N, dtype=torch.bfloat16, device=dev) if mode == L.TC_BF16X3 else None
        ep = L.Epilogue(L.ptr(b), L.ptr(r), L.ptr(toks) if res else None, act, N if res else 0, L.ptr(of), L.ptr(oh), L.ptr(ol), N, DBG)
        def run(i):
            L.call("navc_linear_tc", mode, L.ptr(xs[i % nb]), L.ptr(xl[i % nb]), K, L.ptr(wh), L.ptr(wl), K, m, N, K, ep, L.stream())
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for i in range(reps):
            run(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        fl = 2.0 * m * N * K
        mult = 3 if mode == L.TC_BF16X3 else 1
        print("%-6s %-5s M=%d N=%d K=%d  %7.1f us  useful %6.1f TF/s  issued %6.1f TF/s" % (mname, name, m, N, K, us, fl / us / 1e6, mult * fl / us / 1e6))
