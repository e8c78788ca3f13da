#!/usr/bin/env python
"""BASELINE config 4: NACF inference, all decoding algorithms (mp / ef / l2r, with coarse-grained templates),
q in {1,2,3}, batch 512, one B200.  Prints one JSON line per (paradigm, q)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, cases, navc_b200
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
base = cases.config2()
torch.manual_seed(0)
model = navc_b200.get_model(base).to(dev).eval(); model.set_precision(precision)
feats, category = cases.synth_inputs(base, B)
feats = [f.to(dev) for f in feats]; category = category.to(dev)
for paradigm, qs in (("mp", (1,)), ("ef", (1, 2, 3)), ("l2r", (1, 2, 3))):
    for q in qs:
        opt = dict(base, paradigm=paradigm, q=q, use_ct=True, q_iterations=1)
        tr = navc_b200.Translator(model, opt, device=dev)
        def step():
            with torch.no_grad():
                enc = model.encode(feats=feats)
                return tr.translate_batch(enc, category, None, {})[0]
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n): hyp = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        st = navc_b200.generate.last_stats
        print(json.dumps({"metric": "captions/sec", "paradigm": paradigm, "q": q, "use_ct": True, "batch": B, "precision": precision,
                          "ms_per_batch": round(ms, 2), "value": round(B / ms * 1e3, 1), "decoder_passes": st["passes"],
                          "cuda_graph": st["graph"], "packed_rows": st["packed"], "Smax": st["S"]}), flush=True)
