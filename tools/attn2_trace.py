#!/usr/bin/env python
"""In-kernel timeline of one attention2 launch at the config-2 shape (navc_debug_trace_attn): python tools/attn2_trace.py self|cross"""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L

dev = torch.device("cuda", 0)
L.ensure_init(dev)
kind = sys.argv[1] if len(sys.argv) > 1 else "cross"
B, group, S, E, D, H = 128, 6, 28, 120, 512, 8
N = B * group
gen = torch.Generator().manual_seed(3)
lens = torch.randint(4, 24, (N,), generator=gen).int()
off = torch.zeros(N + 1, dtype=torch.int32); off[1:] = torch.cumsum(lens, 0)
off_d = off.to(dev)
R = N * S
chi = torch.zeros(R, D, dtype=torch.bfloat16, device=dev); clo = torch.zeros_like(chi)
if kind == "self":
    qkv = torch.randn(R, 3 * D, device=dev).to(torch.bfloat16); qkl = qkv * 0.01
    toks = torch.ones(N, S, dtype=torch.int64, device=dev)
    w = L._lib.navc_attention_window(); n_tiles = (R + w - 1) // w
    ts = torch.empty(n_tiles + 1, dtype=torch.int32, device=dev)
    L.call("navc_pack_tiles", L.ptr(off_d), N, L.ptr(ts), n_tiles, L.stream())
    run = lambda: L.call("navc_self_attention_tc_tiles", L.TC_BF16X3, L.ptr(qkv), L.ptr(qkl), 3 * D, L.ptr(toks), L.ptr(off_d), L.ptr(ts), n_tiles,
                         R, N, S, D, H, 0, 0, L.ptr(chi), L.ptr(clo), L.stream())
else:
    q = torch.randn(R, D, device=dev).to(torch.bfloat16); ql = q * 0.01
    kv = torch.randn(B * E, 2 * D, device=dev).to(torch.bfloat16); kvl = kv * 0.01
    run = lambda: L.call("navc_cross_attention_tc_tiles", L.TC_BF16X3, L.ptr(q), L.ptr(ql), D, L.ptr(kv), L.ptr(kvl), 2 * D, L.ptr(off_d), R, N, S, E, D, H,
                         group, L.ptr(chi), L.ptr(clo), L.stream())
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
print("%s: %.1f us per launch (eager, 10 launches), rows %d" % (kind, e0.elapsed_time(e1) * 100, int(off[-1])))
buf = torch.zeros(148 * 3 * 64, dtype=torch.int64, device=dev)
fn = L._lib.navc_debug_trace_attn
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.data_ptr()) == 0
run(); torch.cuda.synchronize(); fn(None)
t = buf.cpu().view(148, 3, 64)
NAMES = {1: "start", 2: "geom", 3: "issue", 4: "qk", 5: "S", 6: "PV", 7: "Srdy", 8: "Sreg", 9: "P", 10: "Ordy", 11: "stored", 12: "Oreg", 13: "lock"}
GHZ = float(os.environ.get("GHZ", "1.9"))
for cta in (0, 77):
    for role in range(3):
        ev = [((x & ((1 << 56) - 1)), (x >> 56) & 0xff) for x in t[cta, role].tolist() if x]
        if not ev:
            continue
        t0 = min(e[0] for r_ in range(3) for e in [((x & ((1 << 56) - 1)), 0) for x in t[cta, r_].tolist() if x])
        print("CTA %3d role %d: " % (cta, role) + " ".join("%s@%.1f" % (NAMES.get(tag, str(tag)), (c - t0) / GHZ / 1e3) for c, tag in ev[:40]))
# distribution over CTAs of the busy span (clock64 is per SM: only differences inside a CTA are meaningful)
spans, items = [], []
for cta in range(148):
    xs = [x & ((1 << 56) - 1) for x in t[cta].flatten().tolist() if x]
    if xs:
        spans.append((max(xs) - min(xs)) / GHZ / 1e3)
        items.append(sum(1 for x in t[cta, 1].tolist() if x and ((x >> 56) & 0xff) == 4) - 1)
q = lambda v, p: sorted(v)[min(len(v) - 1, int(p * len(v)))]
print("per-CTA busy span us: min %.1f median %.1f p90 %.1f max %.1f | items per CTA: min %d max %d"
      % (min(spans), q(spans, .5), q(spans, .9), max(spans), min(items), max(items)))
