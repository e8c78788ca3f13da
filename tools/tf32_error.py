#!/usr/bin/env python
"""Log-probability error of every precision mode on the golden forward fixtures (vs the reference's fp32 values)."""
import glob, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import cases, navc_b200
from oracle import navc_oracle as O
DEV = torch.device("cuda", 0)
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fwd_*.pt"))):
    g = torch.load(path, weights_only=False); opt = g["opt"]
    line = "%-24s max|logp| %5.1f " % (os.path.basename(path), max(b.abs().max().item() for b in g["logprobs"]))
    for prec in ("bf16x3", "tf32", "bf16"):
        m = navc_b200.get_model(opt); m.load_state_dict(cases.synth_state_dict(g["shapes"], g["wseed"], g.get("wscale", 1.0)))
        m.to(DEV).eval(); m.set_precision(prec)
        feats, cat = cases.synth_inputs(opt, g["batch"]); nar = O.is_nar(opt)
        toks = cases.synth_tokens(opt, g["batch"], kind="nar" if nar else "ar"); dis = opt["decoder"] == "BertDecoderDisentangled"
        tgt = [toks["tokens_1"], toks["tokens"]] if (dis and nar) else ([toks["tokens"], toks["tokens"]] if dis else toks["tokens"])
        td = lambda x: [t.to(DEV) for t in x] if isinstance(x, (list, tuple)) else x.to(DEV)
        with torch.no_grad():
            res = m(feats=td(feats), tgt_tokens=td(tgt), category=cat.to(DEV))
        err = max((a.cpu() - b).abs().max().item() for a, b in zip(res["tgt_word_logprobs"], g["logprobs"]))
        line += " %s %.2e" % (prec, err)
    print(line)
