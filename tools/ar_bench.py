#!/usr/bin/env python
"""Autoregressive (ARB) beam-search throughput at the config-2 model size: device-side search with a K/V cache
vs the host-side cross-check implementation.  python tools/ar_bench.py [--batch 128] [--beam 5] [--steps 5]"""
import argparse
import json
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
from navc_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--profile", action="store_true", help="per-kernel device-time breakdown of one batch (device-side search)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    kw = dict(dim_hidden=512, num_hidden_layers_decoder=6, intermediate_size=2048, dim_i=2048, dim_m=2048, n_frames=60,
              max_len=30, vocab_size=10547, beam_size=args.beam, topk=1, beam_alpha=1.0)
    for host in (False, True):
        opt = cases.make_opt("ARB", navc_ar_host_beam=host, **kw)
        torch.manual_seed(0)
        model = navc_b200.get_model(opt).to(dev).eval()
        model.set_precision(args.precision)
        tr = navc_b200.Translator(model, opt, device=dev)
        feats, category = cases.synth_inputs(opt, args.batch)
        feats, category = [f.to(dev) for f in feats], category.to(dev)
        lens = []
        with torch.no_grad():
            for _ in range(2):
                tr.translate_batch(model.encode(feats=feats), category, None, {})
            torch.cuda.synchronize()
            l0 = L.launches
            t0 = time.perf_counter()
            for _ in range(args.steps):
                hyps, _ = tr.translate_batch(model.encode(feats=feats), category, None, {})
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.steps
            lens = [len(h[0]) for h in hyps]
        if args.profile and not host:
            from torch.profiler import profile, ProfilerActivity
            with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
                tr.translate_batch(model.encode(feats=feats), category, None, {})
                torch.cuda.synchronize()
            agg = {}
            for ev in prof.events():
                if ev.device_type == torch.autograd.DeviceType.CUDA:
                    a = agg.setdefault(ev.name[:90], [0, 0.0]); a[0] += 1; a[1] += ev.device_time
            tot = sum(v[1] for v in agg.values())
            print("total device time: %.2f ms over %d kernels" % (tot / 1e3, sum(v[0] for v in agg.values())), file=sys.stderr)
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
                print("%9.1f us %5d x %8.1f us  %5.1f%%  %s" % (v[1], v[0], v[1] / v[0], 100 * v[1] / tot, k), file=sys.stderr)
        print(json.dumps({"metric": "captions/sec (ARB beam search, beam %d, max_len 30)" % args.beam,
                          "impl": "host-side beam, whole-prefix decoder pass per step" if host else "device-side beam, K/V cache",
                          "value": args.batch / dt, "ms_per_batch": dt * 1e3, "batch": args.batch, "dtype": args.precision,
                          "gpu_launches_per_batch": (L.launches - l0) // args.steps,
                          "mean_caption_len": sum(lens) / len(lens)}))


if __name__ == "__main__":
    main()
