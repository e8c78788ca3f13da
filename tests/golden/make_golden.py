"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt

Fixtures hold: the opt dict, the state_dict inventory (names+shapes; weights are re-derived from
a seed by cases.synth_state_dict), and the reference outputs (log-probs, encoder outputs, token
ids for every decode algorithm, loss and selected gradients).  Inputs are re-derived from seeds
by cases.synth_inputs / cases.synth_tokens.  Decode goldens are only kept if the run is free of
ties at every decision boundary -- checked with the oracle's margin bookkeeping.
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)

import cases  # noqa: E402
import refutil  # noqa: E402
from oracle import navc_oracle as O  # noqa: E402


def build(opt, wseed, wscale=1.0):
    model = refutil.ref_get_model(opt)
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = cases.synth_state_dict(shapes, wseed, wscale)
    model.load_state_dict(sd)
    model.eval()
    return model, shapes, sd


def forward_case(name, opt, batch, wseed=7):
    model, shapes, sd = build(opt, wseed)
    feats, category = cases.synth_inputs(opt, batch)
    nar = O.is_nar(opt)
    toks = cases.synth_tokens(opt, batch, kind="nar" if nar else "ar")
    dis = opt["decoder"] == "BertDecoderDisentangled"
    if dis and nar:
        tgt, labels = [toks["tokens_1"], toks["tokens"]], [toks["labels_1"], toks["labels"]]
    elif dis:
        tgt, labels = [toks["tokens"], toks["tokens"]], [toks["labels"], toks["labels"]]
    else:
        tgt, labels = toks["tokens"], toks["labels"]
    with torch.no_grad():
        res = model(feats=[f.clone() for f in feats], tgt_tokens=tgt, category=category)
    out = {"kind": "forward", "opt": opt, "shapes": shapes, "wseed": wseed, "batch": batch,
           "logprobs": [t.clone() for t in res["tgt_word_logprobs"]],
           "enc_output": res["enc_output"].clone(), "enc_hidden": res["enc_hidden"].clone()}
    if "pred_length" in res:
        out["pred_length"] = res["pred_length"].clone()
    # training-mode loss / gradients with dropout disabled (BN uses batch statistics)
    topt = dict(opt, hidden_dropout_prob=0.0, encoder_dropout=0.0)
    tmodel, _, _ = build(topt, wseed)
    tmodel.train()
    res = tmodel(feats=[f.clone() for f in feats], tgt_tokens=tgt, category=category)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss = O.criterion(topt, res, labels, toks.get("length_target") if nar else None)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in tmodel.named_parameters() if p.grad is not None}
    keep = [k for k in grads if any(s in k for s in (
        "Encoder_M.0.bias", "Encoder_I.1.w2.bias", "bn0.weight", "net.3.weight", "position_embeddings",
        "embedding.LayerNorm.bias", "layer.0.attention.self.query.weight", "attend_to_enc_output.self.value.bias",
        "layer.0.intermediate.dense.bias", "layer.1.output.dense.bias"))]
    out.update(loss=loss.detach().clone(), grad_norms={k: g.norm().item() for k, g in grads.items()},
               grads={k: grads[k] for k in keep},
               bn_running={k: v.clone() for k, v in tmodel.state_dict().items() if "running_" in k})
    torch.save(out, os.path.join(HERE, name + ".pt"))
    print("wrote", name, "loss", float(loss.detach()))


def decode_case(name, base_opt, batch, grid, teacher_opt=None, wseed=7, wscale=1.0):
    out = {"kind": "decode", "opt": base_opt, "wseed": wseed, "wscale": wscale, "batch": batch, "runs": []}
    model, shapes, sd = build(base_opt, wseed, wscale)
    out["shapes"] = shapes
    teacher = tsd = None
    if teacher_opt is not None:
        teacher, tshapes, tsd = build(teacher_opt, wseed + 1, wscale)
        out["teacher_opt"], out["teacher_shapes"] = teacher_opt, tshapes
    feats, category = cases.synth_inputs(base_opt, batch)
    for kw in grid:
        opt = dict(base_opt, **kw)
        hyp, _ = refutil.ref_translate(model, opt, feats, category, teacher_model=teacher)
        hyp_o, det = O.translate(sd, opt, feats, category,
                                 teacher=(tsd, teacher_opt) if teacher is not None else None,
                                 return_details=True)
        tie_free = det["min_select_gap"] > 0 and det["min_top2_gap"] > 0 and det["min_candidate_gap"] > 0
        assert torch.equal(hyp, hyp_o), (name, kw)
        assert tie_free, (name, kw, det["min_select_gap"], det["min_top2_gap"], det["min_candidate_gap"])
        out["runs"].append({"kw": kw, "hyp": hyp.clone(), "passes": det["passes"],
                            "min_top2_gap": det["min_top2_gap"], "min_select_gap": det["min_select_gap"],
                            "min_candidate_gap": det["min_candidate_gap"], "beam": det["beam"].clone(),
                            "video_margin": det["video_margin"].clone()})
        print(name, kw, "passes", det["passes"], "gaps %.2e %.2e %.2e" % (
            det["min_top2_gap"], det["min_select_gap"], det["min_candidate_gap"]))
    torch.save(out, os.path.join(HERE, name + ".pt"))


WIDE_SCALE = float(os.environ.get("WIDE_SCALE", "0.4"))   # D = 512: matrices ~N(0, 0.032)


def main():
    assert refutil.reference_available(), "run in the build container (needs /root/reference)"
    only = set(sys.argv[1:])

    def want(name):
        return not only or name in only

    global forward_case, decode_case
    _fwd, _dec = forward_case, decode_case
    forward_case = lambda name, *a, **k: _fwd(name, *a, **k) if want(name) else None
    decode_case = lambda name, *a, **k: _dec(name, *a, **k) if want(name) else None
    forward_case("fwd_config1_nab", cases.config1(), 4)
    forward_case("fwd_small_nacf", cases.small("NACF"), 5)
    forward_case("fwd_small_nacf_ln", cases.small("NACF", with_layernorm=True), 5)
    forward_case("fwd_small_arb", cases.small("ARB"), 5)
    forward_case("fwd_small_nab_plain", cases.small("NAB", with_category=False, enhance_input=0,
                                                    no_encoder_bn=True, hidden_act="relu"), 5)
    grid = [dict(paradigm=p, use_ct=c, q=q) for p in ("mp", "ef", "l2r") for c in (False, True) for q in (1, 2)]
    decode_case("dec_config1_nab", cases.config1(), 4,
                [dict(paradigm=p) for p in ("mp", "ef", "l2r")])
    decode_case("dec_small_nacf", cases.small("NACF"), 6, grid)
    decode_case("dec_small_nacf_teacher", cases.small("NACF"), 6,
                [dict(paradigm="mp", use_ct=True), dict(paradigm="mp", use_ct=True, masking_decision=True),
                 dict(paradigm="ef", use_ct=True), dict(paradigm="l2r", use_ct=False),
                 dict(paradigm="mp", use_ct=False, no_candidate_decision=True)],
                teacher_opt=cases.small("ARB"))
    # ---- dk = 64 (the head size of BASELINE configs 2-5): the cases on which the product runs its tcgen05
    # attention cores, packed rows, the second-level vocabulary packing and (mask-predict) CUDA-graph replay ----
    forward_case("fwd_dk64_nacf", cases.dk64("NACF"), 5)
    forward_case("fwd_dk64_arb", cases.dk64("ARB"), 5)
    decode_case("dec_dk64_nacf", cases.dk64("NACF"), 6, grid)
    decode_case("dec_dk64_nacf_teacher", cases.dk64("NACF"), 6,
                [dict(paradigm="mp", use_ct=True), dict(paradigm="mp", use_ct=True, masking_decision=True),
                 dict(paradigm="ef", use_ct=True, q=2), dict(paradigm="l2r", use_ct=False)],
                teacher_opt=cases.dk64("ARB"))
    wide_grid = [dict(paradigm="mp", use_ct=True), dict(paradigm="mp", use_ct=False), dict(paradigm="mp", use_ct=True, iterations=3),
                 dict(paradigm="ef", use_ct=True, q=2), dict(paradigm="ef", use_ct=False, q=3), dict(paradigm="l2r", use_ct=True, q=2)]
    decode_case("dec_wide_nacf", cases.wide("NACF"), 8, wide_grid, wscale=WIDE_SCALE)
    decode_case("dec_wide_nacf_teacher", cases.wide("NACF"), 8,
                [dict(paradigm="mp", use_ct=True), dict(paradigm="mp", use_ct=False, masking_decision=True)],
                teacher_opt=cases.wide("ARB"), wscale=WIDE_SCALE)


if __name__ == "__main__":
    main()
