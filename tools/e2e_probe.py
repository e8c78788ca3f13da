#!/usr/bin/env python
"""Where do the 100-300 ms outliers of the end-to-end loop come from?  Per-step times of loop variants."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, cases, navc_b200
dev = torch.device("cuda", 0)
opt = cases.config2()
torch.manual_seed(0)
model = navc_b200.get_model(opt).to(dev).eval(); model.set_precision("bf16x3")
tr = navc_b200.Translator(model, opt, device=dev)
B, n_rot = 128, 4
host, devin = [], []
for r in range(n_rot):
    feats, category = cases.synth_inputs(opt, B, seed=1234 + 17 * r)
    host.append(([f.pin_memory() for f in feats], category.pin_memory()))
    devin.append(([f.to(dev) for f in feats], category.to(dev)))
slots = [([torch.empty_like(f) for f in devin[0][0]], torch.empty_like(devin[0][1])) for _ in range(2)]
copy_stream = torch.cuda.Stream()
pin_out = torch.empty((B, 32), dtype=torch.int64).pin_memory()

def compute(feats, category):
    enc = model.encode(feats=feats)
    return tr.translate_batch(enc, category, None, {})[0]

def run(name, fn, n=60):
    with torch.no_grad():
        for i in range(8): fn(i)
        torch.cuda.synchronize()
        ts = []
        for i in range(n):
            t0 = time.perf_counter(); fn(i); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    s = sorted(ts)
    print("%-34s median %.2f  p90 %.2f  max %.1f  n>2x %d" % (name, s[n // 2], s[int(n * .9)], s[-1], sum(t > 2 * s[n // 2] for t in ts)), flush=True)

def resident(i): compute(*devin[i % n_rot])
def resident_cpu(i): compute(*devin[i % n_rot]).cpu()
def h2d_sync(i):
    fh, ch = host[i % n_rot]
    compute([f.to(dev, non_blocking=True) for f in fh], ch.to(dev, non_blocking=True))
def h2d_sync_cpu(i):
    fh, ch = host[i % n_rot]
    compute([f.to(dev, non_blocking=True) for f in fh], ch.to(dev, non_blocking=True)).cpu()
def h2d_slots_cpu(i):
    fh, ch = host[i % n_rot]; fd, cd = slots[i % 2]
    for d, h in zip(fd, fh): d.copy_(h, non_blocking=True)
    cd.copy_(ch, non_blocking=True)
    compute(fd, cd).cpu()
def h2d_slots_pinned_out(i):
    fh, ch = host[i % n_rot]; fd, cd = slots[i % 2]
    for d, h in zip(fd, fh): d.copy_(h, non_blocking=True)
    cd.copy_(ch, non_blocking=True)
    hyp = compute(fd, cd)
    pin_out[:, :hyp.shape[1]].copy_(hyp, non_blocking=True)
for name, fn in (("resident", resident), ("resident + hyp.cpu()", resident_cpu), ("H2D .to() same stream", h2d_sync),
                 ("H2D .to() + hyp.cpu()", h2d_sync_cpu), ("H2D into static slots + hyp.cpu()", h2d_slots_cpu),
                 ("H2D slots + pinned D2H", h2d_slots_pinned_out), ("resident (again)", resident)):
    run(name, fn)
