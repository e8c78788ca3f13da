"""The oracle must reproduce the committed golden vectors (generated from the unmodified reference
by tests/golden/make_golden.py).  CPU only; runs on every machine."""
import glob
import os
import warnings

import pytest
import torch

import cases
from oracle import navc_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD = sorted(glob.glob(os.path.join(GOLDEN, "fwd_*.pt")))
DEC = sorted(glob.glob(os.path.join(GOLDEN, "dec_*.pt")))


def load(path):
    return torch.load(path, weights_only=False)


def fwd_inputs(g):
    opt = g["opt"]
    feats, category = cases.synth_inputs(opt, g["batch"])
    nar = O.is_nar(opt)
    toks = cases.synth_tokens(opt, g["batch"], kind="nar" if nar else "ar")
    dis = opt["decoder"] == "BertDecoderDisentangled"
    if dis and nar:
        tgt, labels = [toks["tokens_1"], toks["tokens"]], [toks["labels_1"], toks["labels"]]
    elif dis:
        tgt, labels = [toks["tokens"], toks["tokens"]], [toks["labels"], toks["labels"]]
    else:
        tgt, labels = toks["tokens"], toks["labels"]
    return feats, category, tgt, labels, (toks.get("length_target") if nar else None)


def test_fixtures_present():
    assert len(FWD) >= 5 and len(DEC) >= 3


@pytest.mark.parametrize("path", FWD, ids=[os.path.basename(p)[:-3] for p in FWD])
def test_forward_golden(path):
    g = load(path)
    opt = g["opt"]
    sd = cases.synth_state_dict(g["shapes"], g["wseed"])
    feats, category, tgt, labels, length_target = fwd_inputs(g)
    with torch.no_grad():
        res = O.model_forward(sd, opt, feats, tgt, category)
    for a, b in zip(res["tgt_word_logprobs"], g["logprobs"]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(res["enc_output"], g["enc_output"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(res["enc_hidden"], g["enc_hidden"], rtol=1e-5, atol=1e-5)
    if "pred_length" in g:
        torch.testing.assert_close(res["pred_length"], g["pred_length"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("path", FWD, ids=[os.path.basename(p)[:-3] for p in FWD])
def test_gradient_golden(path):
    g = load(path)
    opt = dict(g["opt"], hidden_dropout_prob=0.0, encoder_dropout=0.0)
    sd = cases.synth_state_dict(g["shapes"], g["wseed"])
    sd = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    feats, category, tgt, labels, length_target = fwd_inputs(g)
    bn_state = {}
    res = O.model_forward(sd, opt, feats, tgt, category, training=True, bn_state=bn_state)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss = O.criterion(opt, res, labels, length_target)
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["loss"], rtol=1e-5, atol=1e-5)
    for k, gn in g["grad_norms"].items():
        assert abs(sd[k].grad.norm().item() - gn) <= 1e-4 * max(1.0, gn), k
    for k, gr in g["grads"].items():
        torch.testing.assert_close(sd[k].grad, gr, rtol=1e-4, atol=1e-5, msg=k)
    for k, v in g["bn_running"].items():
        torch.testing.assert_close(bn_state[k], v, rtol=1e-5, atol=1e-6, msg=k)


@pytest.mark.parametrize("path", DEC, ids=[os.path.basename(p)[:-3] for p in DEC])
def test_decode_golden(path):
    g = load(path)
    sc = g.get("wscale", 1.0)
    sd = cases.synth_state_dict(g["shapes"], g["wseed"], sc)
    teacher = None
    if "teacher_opt" in g:
        teacher = (cases.synth_state_dict(g["teacher_shapes"], g["wseed"] + 1, sc), g["teacher_opt"])
    feats, category = cases.synth_inputs(g["opt"], g["batch"])
    for run in g["runs"]:
        opt = dict(g["opt"], **run["kw"])
        hyp, det = O.translate(sd, opt, feats, category, teacher=teacher, return_details=True)
        assert det["passes"] == run["passes"], run["kw"]
        assert torch.equal(det["beam"], run["beam"]), run["kw"]
        assert torch.equal(hyp, run["hyp"]), run["kw"]
        if "video_margin" in run:
            torch.testing.assert_close(det["video_margin"], run["video_margin"], rtol=1e-3, atol=1e-6)


def test_tie_break_is_lowest_index_first():
    p = torch.tensor([[0.5, 0.2, 0.2, 0.2, 1.0, 1.0]])
    assert O.k_smallest_mask(p, torch.tensor([2])).tolist() == [[False, True, True, False, False, False]]
    assert O.k_largest_mask(p, torch.tensor([1])).tolist() == [[False, False, False, False, True, False]]
    assert O.k_largest_mask(p, torch.tensor([0])).sum() == 0
    # k is clamped to >= 1 for the smallest selection (algorithms.py:213 max(1, .))
    assert O.k_smallest_mask(p, torch.tensor([0])).sum() == 1


def test_num_mask_truncation_matches_exact_floor():
    # float32 len*ratio truncation equals floor(len*(T-t)/T) for T in {5,6}
    for T in (5, 6):
        for t in range(1, T):
            lens = torch.arange(4, 30)
            k = (lens.float() * (1.0 - (t / T))).long()
            assert torch.equal(k, (lens * (T - t)) // T)
