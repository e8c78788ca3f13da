"""Pin the oracle against the UNMODIFIED reference imported from /root/reference.

Runs only where the reference is mounted (the build container); the committed golden vectors
(tests/test_oracle_golden.py) carry the same pinning to machines without it.
"""
import pytest
import torch

import cases
import refutil
from oracle import navc_oracle as O

pytestmark = pytest.mark.skipif(not refutil.reference_available(), reason="/root/reference not mounted")


def _sd(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


@pytest.mark.parametrize("method,kw", [
    ("NAB", {}),
    ("NACF", {}),
    ("NACF", {"with_layernorm": True}),
    ("ARB", {}),
    ("NAB", {"with_category": False, "enhance_input": 0}),
])
def test_forward_logprobs_match(method, kw):
    opt = cases.small(method, **kw)
    model = refutil.ref_get_model(opt)
    sd = _sd(model)
    feats, category = cases.synth_inputs(opt, 5)
    toks = cases.synth_tokens(opt, 5, kind="nar" if O.is_nar(opt) else "ar")
    tgt = [toks["tokens_1"], toks["tokens"]] if method == "NACF" else toks["tokens"]
    with torch.no_grad():
        ref = model(feats=[f.clone() for f in feats], tgt_tokens=tgt, category=category)
        mine = O.model_forward(sd, opt, feats, tgt, category)
    assert len(ref["tgt_word_logprobs"]) == len(mine["tgt_word_logprobs"])
    for a, b in zip(ref["tgt_word_logprobs"], mine["tgt_word_logprobs"]):
        torch.testing.assert_close(b, a, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(mine["enc_output"], ref["enc_output"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mine["enc_hidden"], ref["enc_hidden"], rtol=1e-5, atol=1e-6)
    if "pred_length" in ref:
        torch.testing.assert_close(mine["pred_length"], ref["pred_length"], rtol=1e-5, atol=1e-6)


def test_config1_logits_match():
    opt = cases.config1()
    model = refutil.ref_get_model(opt)
    sd = _sd(model)
    feats, category = cases.synth_inputs(opt, 4)
    toks = cases.synth_tokens(opt, 4)
    with torch.no_grad():
        ref = model(feats=[f.clone() for f in feats], tgt_tokens=toks["tokens"], category=category)
        mine = O.model_forward(sd, opt, feats, toks["tokens"], category)
    torch.testing.assert_close(mine["tgt_word_logprobs"][0], ref["tgt_word_logprobs"][0], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("paradigm", ["mp", "ef", "l2r"])
@pytest.mark.parametrize("use_ct", [False, True])
@pytest.mark.parametrize("q", [1, 2])
def test_translate_ids_match(paradigm, use_ct, q):
    opt = cases.small("NACF", paradigm=paradigm, use_ct=use_ct, q=q)
    model = refutil.ref_get_model(opt)
    sd = _sd(model)
    feats, category = cases.synth_inputs(opt, 6)
    hyp_ref, _ = refutil.ref_translate(model, opt, feats, category)
    hyp, det = O.translate(sd, opt, feats, category, return_details=True)
    # the comparison is only defined without boundary ties
    assert det["min_select_gap"] > 0 and det["min_top2_gap"] > 0 and det["min_candidate_gap"] > 0
    assert torch.equal(hyp, hyp_ref)


@pytest.mark.parametrize("paradigm", ["mp", "ef"])
@pytest.mark.parametrize("masking_decision", [False, True])
def test_translate_with_teacher(paradigm, masking_decision):
    opt = cases.small("NACF", paradigm=paradigm, use_ct=True, masking_decision=masking_decision)
    topt = cases.small("ARB")
    model = refutil.ref_get_model(opt)
    teacher = refutil.ref_get_model(topt, seed=1)
    feats, category = cases.synth_inputs(opt, 6)
    hyp_ref, _ = refutil.ref_translate(model, opt, feats, category, teacher_model=teacher)
    hyp, det = O.translate(_sd(model), opt, feats, category, teacher=(_sd(teacher), topt), return_details=True)
    assert det["min_select_gap"] > 0 and det["min_candidate_gap"] > 0
    assert torch.equal(hyp, hyp_ref)


def test_config1_translate_all_paradigms():
    for paradigm in ("mp", "ef", "l2r"):
        opt = cases.config1(paradigm=paradigm)
        model = refutil.ref_get_model(opt)
        feats, category = cases.synth_inputs(opt, 4)
        hyp_ref, _ = refutil.ref_translate(model, opt, feats, category)
        hyp = O.translate(_sd(model), opt, feats, category)
        assert torch.equal(hyp, hyp_ref), paradigm


def test_gradients_match():
    """Loss + backward parity with dropout disabled and BatchNorm in train mode."""
    opt = cases.small("NACF", hidden_dropout_prob=0.0, encoder_dropout=0.0)
    model = refutil.ref_get_model(opt)
    model.train()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k)
          for k, v in model.state_dict().items()}
    feats, category = cases.synth_inputs(opt, 5)
    toks = cases.synth_tokens(opt, 5)
    ref = model(feats=[f.clone() for f in feats], tgt_tokens=[toks["tokens_1"], toks["tokens"]], category=category)
    ref["tgt_word_labels"] = [toks["labels_1"], toks["labels"]]
    ref_loss = O.criterion(opt, ref, [toks["labels_1"], toks["labels"]], toks["length_target"])
    ref_loss.backward()
    mine = O.model_forward(sd, opt, feats, [toks["tokens_1"], toks["tokens"]], category, training=True)
    loss = O.criterion(opt, mine, [toks["labels_1"], toks["labels"]], toks["length_target"])
    loss.backward()
    torch.testing.assert_close(loss, ref_loss, rtol=1e-5, atol=1e-6)
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        torch.testing.assert_close(sd[name].grad, p.grad, rtol=1e-4, atol=1e-6, msg=name)


def test_criterion_matches_reference():
    opt = cases.small("NACF", hidden_dropout_prob=0.0, encoder_dropout=0.0)
    opt.update(crit_key=[("tgt_word_logprobs", "tgt_word_labels"), ("pred_length", "tgt_length")],
               crit_name=["Cap Loss", "Length Loss"], crit_scale=[1.0, 1.0])
    model = refutil.ref_get_model(opt)
    feats, category = cases.synth_inputs(opt, 5)
    toks = cases.synth_tokens(opt, 5)
    with torch.no_grad():
        res = model(feats=feats, tgt_tokens=[toks["tokens_1"], toks["tokens"]], category=category)
    res["tgt_word_labels"] = [toks["labels_1"], toks["labels"]]
    res["tgt_length"] = toks["length_target"]
    with refutil.reference_on_path():
        import warnings
        from misc.crit import get_criterion
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            crit = get_criterion(dict(opt))
            crit.reset_loss_recorder()
            ref_loss = crit.get_loss(res)
    mine = O.criterion(opt, res, [toks["labels_1"], toks["labels"]], toks["length_target"])
    torch.testing.assert_close(mine, ref_loss, rtol=1e-6, atol=1e-6)


def test_ar_beam_search_matches_reference():
    """Translator.translate_batch_ARFormer + Beam (models/Translator.py:94-161, models/Beam.py) vs the
    oracle's restatement: same hypotheses, same length-normalised scores."""
    opt = cases.small("ARB", beam_size=3, topk=2, beam_alpha=1.0)
    model = refutil.ref_get_model(opt)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    feats, category = cases.synth_inputs(opt, 4)
    with refutil.reference_on_path():
        from models.Translator import Translator
        tr = Translator(model, dict(opt), device=torch.device("cpu"))
        with torch.no_grad():
            enc = model.encode(feats=[f.clone() for f in feats])
            ref_h, ref_s = tr.translate_batch(enc, category, None, {})
    with torch.no_grad():
        mine_h, mine_s = O.ar_beam_search(sd, opt, O.encode(sd, opt, feats), category)
    assert mine_h == ref_h
    for a, b in zip(mine_s, ref_s):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert abs(x - float(y)) < 1e-4


def test_product_criterion_mirrors_reference_criterion():
    """navc_b200.misc.crit (the criterion API the fused cross-entropy plugs into) against the
    reference's misc/crit.py on the same log-prob tensors: loss, loss records, accuracy / perplexity meters."""
    import warnings
    import navc_b200
    from navc_b200.misc import crit as ncrit
    opt = cases.small("NACF", hidden_dropout_prob=0.0, encoder_dropout=0.0)
    opt.update(crit_key=[("tgt_word_logprobs", "tgt_word_labels"), ("pred_length", "tgt_length")],
               crit_name=["Cap Loss", "Length Loss"], crit_scale=[1.0, 1.0])
    model = refutil.ref_get_model(opt)
    feats, category = cases.synth_inputs(opt, 5)
    toks = cases.synth_tokens(opt, 5)
    with torch.no_grad():
        res = model(feats=feats, tgt_tokens=[toks["tokens_1"], toks["tokens"]], category=category)
    res["tgt_word_labels"] = [toks["labels_1"], toks["labels"]]
    res["tgt_length"] = toks["length_target"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with refutil.reference_on_path():
            from misc.crit import get_criterion
            import copy
            rc = get_criterion(copy.deepcopy(opt))  # (the reference appends the meter names to opt['crit_name'] in place)
            rc.reset_loss_recorder()
            ref_loss = rc.get_loss(res)
            ref_fields = rc.get_fieldsnames()
            ref_names, ref_info = rc.get_loss_info()
        mc = ncrit.get_criterion(copy.deepcopy(opt))
        mc.reset_loss_recorder()
        loss = mc.get_loss(res)
        names, info = mc.get_loss_info()
    torch.testing.assert_close(loss, ref_loss, rtol=1e-6, atol=1e-6)
    assert list(names) == list(ref_names) and mc.get_fieldsnames() == ref_fields
    for a, b in zip(info, ref_info):
        assert abs(a - b) < 1e-4 * max(1.0, abs(b))
