#!/usr/bin/env python
"""Chained out-projection + query projection (navc_linear_chain_tc) against the two separate launches, 20 back to back in a CUDA
graph, device-side row count.  python tools/chain_bench.py [rows]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L
dev = torch.device("cuda", 0); L.ensure_init(dev)
live = int(sys.argv[1]) if len(sys.argv) > 1 else 10553
M, D = 21504, 512
md = L.TC_BF16X3
sp = lambda t: (t.to(torch.bfloat16), (t - t.to(torch.bfloat16).float()).to(torch.bfloat16))
cnt = torch.tensor([live], dtype=torch.int32, device=dev)
xs = [sp(torch.randn(M, D, device=dev)) for _ in range(3)]
rh, rl = sp(torch.randn(M, D, device=dev))
w0h, w0l = sp(torch.randn(D, D, device=dev) / math.sqrt(D)); w1h, w1l = sp(torch.randn(D, D, device=dev) / math.sqrt(D))
b0 = torch.randn(D, device=dev); b1 = torch.randn(D, device=dev)
a_h = torch.empty(M, D, dtype=torch.bfloat16, device=dev); a_l = torch.empty_like(a_h); q_h = torch.empty_like(a_h); q_l = torch.empty_like(a_h)
e0 = L.Epilogue(L.ptr(b0), None, None, 0, D, None, L.ptr(a_h), L.ptr(a_l), D, 0, 1, 0, L.ptr(rh), L.ptr(rl), cnt.data_ptr(), live, 0)
e1 = L.Epilogue(L.ptr(b1), None, None, 0, 0, None, L.ptr(q_h), L.ptr(q_l), D, 0, 1, 0, None, None, cnt.data_ptr(), live, 0)
def separate(i):
    xh, xl = xs[i % 3]
    L.call("navc_linear_tc", md, L.ptr(xh), L.ptr(xl), D, L.ptr(w0h), L.ptr(w0l), D, M, D, D, e0, L.stream())
    L.call("navc_linear_tc", md, L.ptr(a_h), L.ptr(a_l), D, L.ptr(w1h), L.ptr(w1l), D, M, D, D, e1, L.stream())
def chained(i):
    xh, xl = xs[i % 3]
    L.call("navc_linear_chain_tc", md, L.ptr(xh), L.ptr(xl), D, L.ptr(w0h), L.ptr(w0l), D, e0, L.ptr(w1h), L.ptr(w1l), D, e1, M, D, D, L.stream())
for name, fn in (("separate (so + cq)", separate), ("chained", chained)):
    s_ = torch.cuda.Stream()
    with torch.cuda.stream(s_):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s_):
            for i in range(20): fn(i)
        g.replay(); torch.cuda.synchronize()
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0_.record(s_)
        for _ in range(5): g.replay()
        e1_.record(s_); torch.cuda.synchronize()
    print("%-20s rows %d: %.1f us per pair of layers" % (name, live, e0_.elapsed_time(e1_) * 10))
# in-kernel timeline of one chained launch (navc_debug_trace, as tools/gemm2_trace.py)
import ctypes
buf = torch.zeros(148 * 3 * 64, dtype=torch.int64, device=dev)
fn = L._lib.navc_debug_trace
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.data_ptr()) == 0
chained(0); torch.cuda.synchronize(); fn(None)
t = buf.cpu().view(148, 3, 64)
gent = [int(x) for x in t[:, 1, 63].tolist() if x]; gext = [int(x) for x in t[:, 1, 62].tolist() if x]
g0 = min(gent)
print("chained grid wall clock: last exit %.2f us after the first entry (exits: min %.2f median %.2f)" % ((max(gext) - g0) / 1e3, (min(gext) - g0) / 1e3, (sorted(gext)[len(gext) // 2] - g0) / 1e3))
t[:, 1, 62:] = 0
NAMES = {1: "prologue", 2: "ld0", 4: "accfree", 5: "op0", 6: "opN", 7: "accrdy", 8: "epi", 9: "drain", 60: "ENTRY", 61: "EXIT"}
for cta in (0, 40, 100, 147):
    ev = []
    for role in (0, 1, 2):
        for x in t[cta, role].tolist():
            if x:
                ev.append((x & ((1 << 56) - 1), (x >> 56) & 0xff))
    ev.sort(); t0 = ev[0][0]
    print("CTA %3d: " % cta + " ".join("%s@%.1f" % (NAMES[tag], (c - t0) / 1.9e3) for c, tag in ev if tag in NAMES))
