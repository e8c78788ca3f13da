"""The drop-in boundary exercised from the REFERENCE's side.

(A) a navc model object handed to the reference's OWN ``models.Translator.Translator`` and ``decoding.generate``
    (unmodified reference code from baseline/_ref or /root/reference): the reference's refinement loop calls
    ``model.prepare_inputs_for_decoder``, ``model.decoder(tgt, **inputs, output_attentions=True)`` and
    ``model.tgt_word_prj(hidden)`` directly (decoding/algorithms.py:143-149, na_generate.py:53-64) -- the ids must be
    the committed goldens (which the same reference produced with its own model);
(B) the ``sys.modules`` alias: after ``sys.modules['models'] = navc_b200.models`` the reference's import lines
    (train.py:81, misc/utils.py:7, misc/run.py:115) resolve to the navc factory / Translator.
"""
import os
import sys

import pytest
import torch

import cases
import navc_b200
import refutil

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not refutil.reference_available(), reason="reference not installed (tools/install_reference.py)")


def _navc_model(opt, shapes, wseed, precision, wscale=1.0):
    model = navc_b200.get_model(opt)
    model.load_state_dict(cases.synth_state_dict(shapes, wseed, wscale))
    model.to(DEV).eval()
    model.set_precision(precision)
    return model


@needs_ref
@pytest.mark.parametrize("fixture", ["dec_small_nacf", "dec_dk64_nacf", "dec_config1_nab"])
def test_reference_translator_and_generate_drive_a_navc_model(fixture):
    g = torch.load(os.path.join(GOLDEN, fixture + ".pt"), weights_only=False)
    model = _navc_model(g["opt"], g["shapes"], g["wseed"], "fp32", g.get("wscale", 1.0))
    feats, category = cases.synth_inputs(g["opt"], g["batch"])
    feats, category = [f.to(DEV) for f in feats], category.to(DEV)
    vocab = {i: "w%d" % i for i in range(g["opt"]["vocab_size"])}
    launches0 = navc_b200._lib.launches
    problems = []
    with refutil.reference_on_path():
        from models.Translator import Translator as RefTranslator   # the reference's class, unmodified
        import decoding as ref_decoding
        assert os.path.realpath(ref_decoding.__file__).startswith(os.path.realpath(refutil.REF_ROOT))
        for run in g["runs"]:
            opt = dict(g["opt"], **run["kw"])
            tr = RefTranslator(model=model, opt=opt, device=DEV, teacher_model=None, dict_mapping={})
            with torch.no_grad():
                enc = model.encode(feats=feats)
                hyp, _ = tr.translate_batch(enc, category, None, vocab)
            margin = min(run["min_top2_gap"], run["min_select_gap"], run["min_candidate_gap"])
            if not torch.equal(hyp.cpu(), run["hyp"]):
                if margin > 2e-5:
                    problems.append((run["kw"], "ids differ, margin %.2e" % margin))
                else:
                    print("NOTE sub-margin decision (%.2e) flipped for %s" % (margin, run["kw"]))
    assert navc_b200._lib.launches > launches0          # the decoder / projection calls ran on the navc kernels
    assert not problems, problems


@needs_ref
def test_reference_generate_with_navc_teacher():
    """Same through ``decoding.generate`` with an ARB teacher (scoring_by_teacher, algorithms.py:175-204)."""
    g = torch.load(os.path.join(GOLDEN, "dec_small_nacf_teacher.pt"), weights_only=False)
    model = _navc_model(g["opt"], g["shapes"], g["wseed"], "fp32")
    teacher = _navc_model(g["teacher_opt"], g["teacher_shapes"], g["wseed"] + 1, "fp32")
    feats, category = cases.synth_inputs(g["opt"], g["batch"])
    feats, category = [f.to(DEV) for f in feats], category.to(DEV)
    problems = []
    with refutil.reference_on_path():
        from decoding import generate as ref_generate
        for run in g["runs"][:2]:
            opt = dict(g["opt"], **run["kw"])
            with torch.no_grad():
                enc, t_enc = model.encode(feats=feats), teacher.encode(feats=feats)
                hyp, _ = ref_generate(opt=opt, model=model, teacher_model=teacher, encoder_outputs=enc,
                                      teacher_encoder_outputs=t_enc, category=category, tgt_tokens=None, tgt_vocab={},
                                      dict_mapping={}, length_bias=0)
            margin = min(run["min_top2_gap"], run["min_select_gap"], run["min_candidate_gap"])
            if not torch.equal(hyp.cpu(), run["hyp"]) and margin > 2e-5:
                problems.append((run["kw"], "ids differ, margin %.2e" % margin))
    assert not problems, problems


def test_sys_modules_alias_resolves_reference_imports():
    saved = {k: sys.modules.get(k) for k in ("models", "models.Translator", "decoding")}
    try:
        sys.modules["models"] = navc_b200.models
        sys.modules["models.Translator"] = navc_b200.models.Translator
        sys.modules["decoding"] = navc_b200.decoding
        from models import get_model                     # train.py:81, misc/utils.py:7
        from models.Translator import Translator         # misc/run.py:19
        from decoding import generate                    # models/Translator.py:166
        assert get_model is navc_b200.get_model and generate is navc_b200.generate
        opt = cases.config1()
        model = get_model(opt).to(DEV)
        feats, category = cases.synth_inputs(opt, 4)
        with torch.no_grad():
            enc = model.encode(feats=[f.to(DEV) for f in feats])
            hyp, _ = Translator(model, opt, device=DEV).translate_batch(enc, category.to(DEV), None, {})
        assert hyp.shape[0] == 4 and hyp.dtype == torch.int64
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
