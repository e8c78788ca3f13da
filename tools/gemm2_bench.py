#!/usr/bin/env python
"""Per-shape timing of the decoder-layer GEMMs through navc_linear_tc (pair epilogue) at the config-2 row count, CUDA
events, inputs rotating over 3 buffers.   python tools/gemm2_bench.py [M] [dbg ...]
dbg (navc_epilogue_t.reserved, profiling aids of gemm2_tc.cu): 0 normal, 11 epilogue = barrier handshake only,
12 no stores, 13 no residual loads, 14 one MMA per tile, 15 no operand loads, 7 no tail split, 128 / 256 forced width."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L

dev = torch.device("cuda", 0)
L.ensure_init(dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10478
dbgs = [int(x) for x in sys.argv[2:]] or [0]
shapes = [("qkv", 1536, 512, 0, False), ("so", 512, 512, 0, True), ("cq", 512, 512, 0, False), ("f1", 2048, 512, 1, False),
          ("f2", 512, 2048, 0, True)]
modes = [(L.TC_BF16X3, "bf16x3")] + ([(L.TC_BF16, "bf16")] if os.environ.get("BF16") else [])
for mode, mname in modes:
    x3 = mode == L.TC_BF16X3
    for name, N, K, act, res in shapes:
        nb = 3
        xs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nb)]
        xl = [torch.randn(M, K, device=dev).to(torch.bfloat16) * 0.01 for _ in range(nb)]
        wh = torch.randn(N, K, device=dev).to(torch.bfloat16); wl = wh * 0.01
        b = torch.randn(N, device=dev)
        rh = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
        rl = (rh * 0.01) if res else None
        toks = torch.ones(M, dtype=torch.int64, device=dev)
        oh = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ol = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if x3 else None
        line = "%-6s %-4s N=%4d K=%4d " % (mname, name, N, K)
        for dbg in dbgs:
            ep = L.Epilogue(L.ptr(b), None, L.ptr(toks) if res else None, act, N if res else 0, None, L.ptr(oh), L.ptr(ol), N, dbg, 1, 0,
                            L.ptr(rh), L.ptr(rl) if x3 else None, None, 0, 0)
            def run(i):
                L.call("navc_linear_tc", mode, L.ptr(xs[i % nb]), L.ptr(xl[i % nb]) if x3 else None, K, L.ptr(wh), L.ptr(wl) if x3 else None, K,
                       M, N, K, ep, L.stream())
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            # device time per launch: 30 launches captured in one CUDA graph (no host launch cost between them), replayed
            reps = 30
            st = torch.cuda.Stream()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st):
                with torch.cuda.graph(graph, stream=st):
                    for i in range(reps):
                        run(i)
                graph.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                graph.replay()
                graph.replay()
                e1.record(st)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (2 * reps)
            line += " dbg%-3d %6.1f us" % (dbg, us)
        print(line)
