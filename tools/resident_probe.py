#!/usr/bin/env python
"""Where do the 100 ms steps of bench.py's resident loop come from?  Replays that loop with host timers around the
calls of a step and prints the slow steps (GPU-side step time from CUDA events + host time of every call)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import cases  # noqa: E402
import navc_b200  # noqa: E402
from navc_b200.decoding import na_generate  # noqa: E402


def main():
    sync_each = "--sync" in sys.argv
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    opt = cases.config2()
    torch.manual_seed(0)
    model = navc_b200.get_model(opt).to(dev).eval()
    model.set_precision("bf16x3")
    tr = navc_b200.Translator(model, opt, device=dev)
    B, n_rot = 128, 4
    devin = []
    for r in range(n_rot):
        feats, category = cases.synth_inputs(opt, B, seed=1234 + 17 * r)
        devin.append(([f.to(dev) for f in feats], category.to(dev)))
    # host timers on the pieces of generate()
    T = {}
    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter()
            out = fn(*a, **k)
            T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return out
        return w
    na_generate._DecodeGraph.replay = timed("replay", na_generate._DecodeGraph.replay)
    na_generate._DecodeGraph.load = timed("load", na_generate._DecodeGraph.load)
    orig_tolist = torch.Tensor.tolist
    torch.Tensor.tolist = timed("tolist", orig_tolist)
    enc_fn = timed("encode", model.encode)
    from navc_b200 import _lib as L
    import gc
    orig_call = L.call
    def call(name, *a):
        t0 = time.perf_counter()
        orig_call(name, *a)
        dt = (time.perf_counter() - t0) * 1e3
        if dt > 1.0:
            T["call:" + name] = T.get("call:" + name, 0.0) + dt
    L.call = call
    import navc_b200.engine as E
    E.L.call = call
    orig_empty = torch.empty
    def empty(*a, **k):
        t0 = time.perf_counter()
        out = orig_empty(*a, **k)
        dt = (time.perf_counter() - t0) * 1e3
        if dt > 1.0:
            T["empty"] = T.get("empty", 0.0) + dt
        return out
    torch.empty = empty
    gc.callbacks.append(lambda phase, info: T.__setitem__("gc_gen%d_%s" % (info["generation"], phase), time.perf_counter() * 1e3))
    if "--freeze" in sys.argv:
        gc.collect(); gc.freeze()

    def step(i):
        feats, category = devin[i % n_rot]
        enc = enc_fn(feats=feats)
        hyp, _ = tr.translate_batch(enc, category, None, {})
        return hyp

    with torch.no_grad():
        for i in range(12):
            step(i)
        torch.cuda.synchronize()
        n = 120
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        host, parts = [], []
        marks[0].record()
        for i in range(n):
            T.clear()
            t0 = time.perf_counter()
            step(i)
            if sync_each:
                torch.cuda.synchronize()
            host.append((time.perf_counter() - t0) * 1e3)
            parts.append(dict(T))
            marks[i + 1].record()
        torch.cuda.synchronize()
    gpu = [marks[i].elapsed_time(marks[i + 1]) for i in range(n)]
    med = sorted(gpu)[n // 2]
    print("sync_each=%s median gpu step %.2f ms, max %.2f; slow steps (> 1.5x median):" % (sync_each, med, max(gpu)))
    for i in range(n):
        if gpu[i] > 1.5 * med or host[i] > 1.5 * med:
            print("  step %3d gpu %.2f host %.2f  " % (i, gpu[i], host[i]) + " ".join("%s=%.2f" % kv for kv in sorted(parts[i].items())))


if __name__ == "__main__":
    main()
