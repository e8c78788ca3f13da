#!/usr/bin/env python
"""In-kernel timeline of one navc_linear_tc launch (gemm2_tc.cu, navc_debug_trace): per CTA the cycle stamps of the
producer / MMA / epilogue roles.   python tools/gemm2_trace.py N K [M] [res]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L

dev = torch.device("cuda", 0)
L.ensure_init(dev)
N, K = int(sys.argv[1]), int(sys.argv[2])
M = int(sys.argv[3]) if len(sys.argv) > 3 else 10478
res = len(sys.argv) > 4 and sys.argv[4] == "res"
xh = torch.randn(M, K, device=dev).to(torch.bfloat16); xl = xh * 0.01
wh = torch.randn(N, K, device=dev).to(torch.bfloat16); wl = wh * 0.01
b = torch.randn(N, device=dev)
rh = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
rl = (rh * 0.01) if res else None
oh = torch.empty(M, N, dtype=torch.bfloat16, device=dev); ol = torch.empty_like(oh)
ep = L.Epilogue(L.ptr(b), None, None, 0, N if res else 0, None, L.ptr(oh), L.ptr(ol), N, 0, 1, 0, L.ptr(rh), L.ptr(rl), None, 0, 0)
def run():
    L.call("navc_linear_tc", L.TC_BF16X3, L.ptr(xh), L.ptr(xl), K, L.ptr(wh), L.ptr(wl), K, M, N, K, ep, L.stream())
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(148 * 3 * 64, dtype=torch.int64, device=dev)
fn = L._lib.navc_debug_trace
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
fn(None)
print("kernel %.1f us (events, eager, traced)" % (e0.elapsed_time(e1) * 1e3))
g = torch.cuda.CUDAGraph()
s_ = torch.cuda.Stream()
with torch.cuda.stream(s_):
    run(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s_):
        for _ in range(20):
            run()
    g.replay(); torch.cuda.synchronize()
    e0.record(s_); g.replay(); e1.record(s_); torch.cuda.synchronize()
print("kernel %.1f us per launch (20 back to back in a graph, untraced)" % (e0.elapsed_time(e1) * 1e3 / 20))
t = buf.cpu().view(148, 3, 64)
# wall-clock (globaltimer, ns) entry / exit of every CTA: launch skew and the span of the whole grid
gent = [int(x) for x in t[:, 1, 63].tolist() if x]; gext = [int(x) for x in t[:, 1, 62].tolist() if x]
if gent and gext:
    g0 = min(gent)
    print("grid wall clock: CTA entries spread over %.2f us, last exit %.2f us after the first entry (exits: min %.2f median %.2f)"
          % ((max(gent) - g0) / 1e3, (max(gext) - g0) / 1e3, (min(gext) - g0) / 1e3, (sorted(gext)[len(gext) // 2] - g0) / 1e3))
t[:, 1, 62:] = 0
NAMES = {**{20 + c: "ld%d" % c for c in range(4)}, **{30 + c: "cmp%d" % c for c in range(4)}, **{40 + c: "free%d" % c for c in range(4)},
         **{50 + c: "st%d" % c for c in range(4)}}
NAMES.update({1: "prologue", 2: "ld0", 3: "ldN", 4: "accfree", 5: "op0", 6: "opN", 7: "accrdy", 8: "epi", 9: "drain", 60: "ENTRY", 61: "EXIT"})
GHZ = float(os.environ.get("GHZ", "1.9"))
ROLES = [int(x) for x in os.environ.get("ROLES", "0,1,2").split(",")]
for cta in (0, 1, 75, 146):
    ev = []
    for role in ROLES:
        for x in t[cta, role].tolist():
            if x:
                ev.append((x & ((1 << 56) - 1), (x >> 56) & 0xff, role))
    if not ev:
        continue
    ev.sort()
    t0 = ev[0][0]
    print("CTA %3d: " % cta + "  ".join("%s@%.1f" % (NAMES.get(tag, str(tag)), (c - t0) / GHZ / 1e3) for c, tag, role in ev))
# spread of the end stamps over CTAs (relative to each CTA's own first stamp)
ends = []
for cta in range(148):
    xs = [x & ((1 << 56) - 1) for x in t[cta].flatten().tolist() if x]
    if xs:
        ends.append((max(xs) - min(xs)) / GHZ / 1e3)
print("per-CTA busy span us: min %.1f median %.1f max %.1f (n=%d)" % (min(ends), sorted(ends)[len(ends) // 2], max(ends), len(ends)))
