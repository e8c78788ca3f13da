"""Model factory with the reference signature: ``get_model(opt: dict) -> nn.Module``
(contract: reference models/__init__.py:64-94; opt keys listed in SURVEY.md section 8b)."""
import torch.nn as nn

from ..config import Constants
from . import Decoder, Encoder, Predictor
from .modules import Joint_Representaion_Learner
from .seq2seq import Seq2Seq, VocabProjection


def _pick(module, name, kind):
    if name not in module.__all__:
        raise ValueError("We can not find %s in models/%s.py" % (name, kind))
    return getattr(module, name)


def get_auxiliary_task_predictor(opt):
    supported = [n[len("Predictor_"):] for n in dir(Predictor) if n.startswith("Predictor_")]
    layers = [getattr(Predictor, "Predictor_%s" % c)(opt, key_name=Constants.mapping[c][0])
              for c in opt["crit"] if c in supported]
    return Predictor.Auxiliary_Task_Predictor(layers) if layers else None


def get_model(opt):
    dims = {"i": opt["dim_i"], "m": opt["dim_m"], "a": opt["dim_a"], "o": opt["dim_o"]}
    for ch in opt["modality"].lower():
        assert ch in dims
    assert not opt.get("use_preEncoder", False)
    # construction order = parameter initialisation order of the reference factory
    encoder = _pick(Encoder, opt["encoder"], "Encoder")(opt)
    jrl = None
    if not opt.get("no_joint_representation_learner", False):
        jrl = Joint_Representaion_Learner([opt["dim_hidden"]] * len(opt["modality"]), opt)
    aux = get_auxiliary_task_predictor(opt)
    decoder = _pick(Decoder, opt["decoder"], "Decoder")(opt)
    prj = VocabProjection(opt["dim_hidden"], opt["vocab_size"], bias=False)
    return Seq2Seq(opt=opt, preEncoder=None, encoder=encoder, joint_representation_learner=jrl,
                   auxiliary_task_predictor=aux, decoder=decoder, tgt_word_prj=prj)
