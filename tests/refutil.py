"""Helpers to import the UNMODIFIED reference: from /root/reference in the build container, else from
baseline/_ref/ (the git-ignored, digest-checked install made by tools/install_reference.py, which travels to
the GPU box with gpurun).

The reference is never part of the product: it is imported only to pin the oracle, to generate golden vectors,
as the CPU baseline / reference arm of bench.py, and by the boundary test that drives the reference's own
Translator over a navc model.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")  # tools/install_reference.py (unmodified copy, git-ignored)


def _pick_root():
    env = os.environ.get("NAVC_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", INSTALLED):
        if os.path.isfile(os.path.join(cand, "models", "seq2seq.py")):
            return cand
    return "/root/reference"


REF_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "seq2seq.py"))


@contextlib.contextmanager
def reference_on_path():
    """Temporarily put the reference first on sys.path and isolate its top-level module names
    (models, decoding, misc, config) so they never collide with the product package."""
    names = ("models", "decoding", "misc", "config")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in names}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k.split(".")[0] in names]:
            del sys.modules[k]
        sys.modules.update(saved)


def ref_get_model(opt, seed=0):
    import torch
    with reference_on_path():
        from models import get_model
        torch.manual_seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            model = get_model(dict(opt))
    model.eval()
    return model


def ref_translate(model, opt, feats, category, teacher_model=None, vocab=None):
    """model.encode + Translator.translate_batch exactly as misc/run.py:130-141 drives them."""
    import torch
    with reference_on_path():
        from models.Translator import Translator
        import decoding  # noqa: F401  (Translator imports it lazily; keep it resolvable)
        tr = Translator(model, dict(opt), device=torch.device("cpu"), teacher_model=teacher_model)
        with torch.no_grad():
            enc = model.encode(feats=[f.clone() for f in feats])
            t_enc = teacher_model.encode(feats=[f.clone() for f in feats]) if teacher_model is not None else None
            vocab = vocab or {i: "w%d" % i for i in range(opt["vocab_size"])}
            hyp, _ = tr.translate_batch(enc, category, None, vocab, teacher_encoder_outputs=t_enc)
    return hyp, enc
