#!/usr/bin/env python
"""BASELINE config 4: NACF inference, all decoding algorithms (mp / ef / l2r), length-parallel q in {1,2,3}, batch 512,
one B200.  Prints one JSON line per (paradigm, use_ct, q).

Both template settings are timed: with random-init weights the coarse-grained template pass (use_ct=True, the NACF
default) leaves no <mask> behind, so easy-first / left-to-right finish after 2 decoder passes -- the ceil(S/q) loop that
defines those algorithms only runs with use_ct=False (q_iterations follows translate.py:142-143: 1 with templates, else 0)."""
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
for use_ct, paradigm, qs in ((True, "mp", (1,)), (True, "ef", (1, 2, 3)), (True, "l2r", (1, 2, 3)),
                             (False, "mp", (1,)), (False, "ef", (1, 2, 3)), (False, "l2r", (1, 2, 3))):
    for q in qs:
        opt = dict(base, paradigm=paradigm, q=q, use_ct=use_ct, q_iterations=1 if use_ct else 0)
        tr = navc_b200.Translator(model, opt, device=dev)
        def step():
            with torch.no_grad():
                enc = model.encode(feats=feats)
                return tr.translate_batch(enc, category, None, {})[0]
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5 if use_ct or paradigm == "mp" else 2
        e0.record()
        for _ in range(n): hyp = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        st = navc_b200.generate.last_stats
        print(json.dumps({"metric": "captions/sec", "paradigm": paradigm, "q": q, "use_ct": use_ct, "batch": B, "precision": precision,
                          "ms_per_batch": round(ms, 2), "value": round(B / ms * 1e3, 1), "decoder_passes": st["passes"],
                          "cuda_graph": st["graph"], "packed_rows": st["packed"], "Smax": st["S"]}), flush=True)
