#!/usr/bin/env python
"""Timing + in-kernel timeline of the attention backward kernels at the training shape (packed rows).
python tools/attn_bwd_trace.py [videos]   -- tcgen05 (attention_bwd_tc.cu) against the CUDA-core kernel (backward.cu)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import navc_b200
from navc_b200 import _lib as L

dev = torch.device("cuda", 0)
L.ensure_init(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S, E, D, H = 30, 120, 512, 8
g = torch.Generator().manual_seed(0)
lens = torch.randint(4, 29, (N,), generator=g, dtype=torch.int32)
seq_off = torch.zeros(N + 1, dtype=torch.int32); seq_off[1:] = torch.cumsum(lens, 0)
R = int(seq_off[-1]); seq_off = seq_off.to(dev)
toks = torch.ones(N, S, dtype=torch.int64, device=dev)
q = torch.randn(R, D, device=dev); kv = torch.randn(N * E, 2 * D, device=dev); qkv = torch.randn(R, 3 * D, device=dev)
d_ctx = torch.randn(R, D, device=dev); ctx = torch.randn(R, D, device=dev)
d_q = torch.empty(R, D, device=dev); d_kv = torch.empty(N * E, 2 * D, device=dev); d_qkv = torch.empty(R, 3 * D, device=dev)
P = L.ptr
def cross_tc(): L.call("navc_cross_attention_bwd_tc", L.TC_BF16X3, P(q), D, P(kv), 2 * D, P(seq_off), N, S, E, D, H, P(d_ctx), P(ctx), P(d_q), D, P(d_kv), 2 * D, L.stream())
def cross_cc(): L.call("navc_cross_attention_bwd_packed", P(q), D, P(kv), 2 * D, P(seq_off), N, S, E, D, H, P(d_ctx), P(d_q), D, P(d_kv), 2 * D, L.stream())
def self_tc(): L.call("navc_self_attention_bwd_tc", L.TC_BF16X3, P(qkv), 3 * D, P(seq_off), N, S, D, H, 0, 0, P(d_ctx), P(ctx), P(d_qkv), L.stream())
def self_cc(): L.call("navc_self_attention_bwd_packed", P(qkv), 3 * D, P(toks), P(seq_off), N, S, D, H, 0, 0, P(d_ctx), P(d_qkv), L.stream())
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("videos %d rows %d" % (N, R))
for name, fn in (("cross tc", cross_tc), ("cross cuda-core", cross_cc), ("self tc", self_tc), ("self cuda-core", self_cc)):
    print("%-16s %8.1f us" % (name, timeit(fn)))
fn_t = L._lib.navc_debug_trace_attn_bwd
fn_t.argtypes = [ctypes.c_void_p]
GHZ = float(os.environ.get("GHZ", "1.9"))
NAMES = ["start", "loads issued", "operands in smem", "S,dP ready", "P,dS written", "dV,dK,dQ ready", "staged", "stores issued"]
for name, fn in (("cross", cross_tc), ("self", self_tc)):
    buf = torch.zeros(N * H * 16, dtype=torch.int64, device=dev)
    assert fn_t(buf.data_ptr()) == 0
    fn(); torch.cuda.synchronize(); fn_t(None)
    t = buf.cpu().view(N * H, 16)
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    spans = []
    for cta in (0, 1, 150, 400, 3000, N * H - 1):
        xs = [int(x) for x in t[cta].tolist() if x]
        print("%s CTA %5d: start@%.1f us  " % (name, cta, (xs[0] - t0) / GHZ / 1e3) + "  ".join("%s +%.2f" % (NAMES[i], (xs[i] - xs[i - 1]) / GHZ / 1e3) for i in range(1, len(xs))))
    tot = [(int(r[7]) - int(r[0])) / GHZ / 1e3 for r in t.tolist() if r[7]]
    import statistics
    print("%s: per-CTA span median %.2f us, mean of phases (us): " % (name, statistics.median(tot)) +
          "  ".join("%s %.2f" % (NAMES[i], statistics.mean((int(r[i]) - int(r[i - 1])) / GHZ / 1e3 for r in t.tolist() if r[7])) for i in range(1, 8)))
