#!/usr/bin/env python
"""Per-kernel device-time breakdown of one bench step (torch.profiler/CUPTI; not a bench number)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import cases
import navc_b200
from torch.profiler import profile, ProfilerActivity

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
# "range": only one step is visible to `ncu --profile-from-start off` (cudaProfilerStart/Stop around it), then exit
RANGE = len(sys.argv) > 3 and sys.argv[3] == "range"
dev = torch.device("cuda", 0)
opt = cases.config2()
torch.manual_seed(0)
model = navc_b200.get_model(opt).to(dev).eval()
model.set_precision(precision)
tr = navc_b200.Translator(model, opt, device=dev)
feats, category = cases.synth_inputs(opt, B)
feats = [f.to(dev) for f in feats]; category = category.to(dev)
def step():
    enc = model.encode(feats=feats)
    return tr.translate_batch(enc, category, None, {})[0]
with torch.no_grad():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if RANGE:
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        sys.exit(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    print("step ms (events): %.2f" % e0.elapsed_time(e1))
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg.setdefault(ev.name[:90], [0, 0.0]); a[0] += 1; a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print("total device time: %.2f ms over %d kernels" % (tot / 1e3, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.1f us %5d x %8.1f us  %5.1f%%  %s" % (v[1], v[0], v[1] / v[0], 100 * v[1] / tot, k))
