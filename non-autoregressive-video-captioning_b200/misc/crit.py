"""Loss / meters with the reference's criterion API (contract: reference misc/crit.py:10-251).

``get_criterion(opt)`` -> ``Criterion`` with ``reset_loss_recorder / get_loss(results) /
get_loss_info / get_fieldsnames``.  ``LanguageGeneration`` accepts either the reference's
``tgt_word_logprobs`` tensors [B, S, V] or, when the model runs with ``opt['navc_fused_ce']``,
``LazyLogProbs`` entries: then projection + log-softmax + masked NLL are ONE fused autograd node
(no [B, S, V] tensor is ever written) and the word-accuracy /
perplexity meters are fed from the same kernel's per-token statistics.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..config import Constants


class AverageMeter:
    """reference misc/logger.py:51-70."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1, multiply=True):
        self.val = val
        self.sum += val * n if multiply else val
        self.count += n
        self.avg = self.sum / self.count if self.count else 0


class CritBase(nn.Module):
    def __init__(self, crit_name, weights=1.0, batch_mean=True):
        super().__init__()
        assert crit_name in Constants.mapping
        self.keys = Constants.mapping[crit_name]
        self.weights = weights
        self.batch_mean = batch_mean

    def _step(self, *inputs):
        raise NotImplementedError()

    def forward(self, kwargs):
        src1, src2, *others = [kwargs[k] for k in self.keys]
        src1 = src1 if isinstance(src1, list) else [src1]
        src2 = src2 if isinstance(src2, list) else [src2] * len(src1)
        assert len(src1) == len(src2)
        weights = self.weights if isinstance(self.weights, list) else [self.weights] * len(src1)
        assert len(src1) == len(weights)
        denom = src1[0].size(0) if self.batch_mean else 1.0
        loss = None
        for i, (w, a, b) in enumerate(zip(weights, src1, src2)):
            term = w * self._step(i, a, b, *others) / denom
            loss = term if loss is None else loss + term
        return loss, denom


class LanguageGeneration(CritBase):
    def __init__(self, opt, crit_name, weights=1.0, batch_mean=True):
        vw = opt.get("visual_word_generation", False)
        if vw:
            weights = opt.get("nv_weights", [0.8, 1.0])
        super().__init__(crit_name, weights, batch_mean)
        self.ignore_index = Constants.PAD
        self.num_word_acc = 2 if vw else 1
        self.visual_word_generation = vw
        self.reset_recorder()

    def _step(self, index, logprobs, labels, *others):
        assert not others
        assert logprobs.size(1) == labels.size(1)
        if hasattr(logprobs, "nll_sum"):  # fused path: LazyLogProbs
            loss = logprobs.nll_sum(labels)
            nll, pred = logprobs.stats
            self._meters(index, pred.long(), -nll, labels)
            return loss
        pred = logprobs.max(-1)[1]
        tok_lp = logprobs.gather(2, labels.unsqueeze(2)).squeeze(2)
        self._meters(index, pred, tok_lp, labels)
        flat = logprobs.contiguous().view(-1, logprobs.size(2))
        lab = labels.contiguous().view(-1)
        nll = nn.functional.nll_loss(flat, lab, reduction="none")
        return torch.sum(nll * lab.ne(self.ignore_index).float())

    def _meters(self, index, pred, tok_logprob, labels):
        ind = labels.ne(Constants.PAD)
        if index == 0 and self.visual_word_generation:
            ind = ind & labels.ne(Constants.MASK)
        n = int(ind.sum().item())
        if n:
            self.word_acc_recorder[index].update(int((pred[ind] == labels[ind]).sum().item()), n, multiply=False)
        if index == 0 and self.visual_word_generation:
            return  # perplexity is reported for the caption pass only (crit.py:104-105)
        mask = labels.ne(Constants.PAD)
        num_words = float(mask.sum().item())
        if num_words:
            self.perplexity_recorder.update((-(tok_logprob * mask).sum() / num_words).item(), num_words)

    def get_fieldsnames(self):
        return ["Word Acc%d" % i for i in range(self.num_word_acc)] + ["Perplexity"]

    def get_info(self):
        return self.get_fieldsnames(), [m.avg for m in self.word_acc_recorder] + [math.exp(self.perplexity_recorder.avg)]

    def reset_recorder(self):
        self.word_acc_recorder = [AverageMeter() for _ in range(self.num_word_acc)]
        self.perplexity_recorder = AverageMeter()


class Criterion:
    def __init__(self, crit_objects, keys, names, scales, summarywriter=None):
        assert len(crit_objects) == len(keys) == len(names) == len(scales)
        self.crit_objects, self.keys, self.names, self.scales = crit_objects, keys, list(names), scales
        self.num_loss = len(crit_objects)
        self.summarywriter = summarywriter
        self.n_current_round = 0
        self.reset_loss_recorder()

    def reset_loss_recorder(self):
        self.loss_recorder = [AverageMeter() for _ in range(self.num_loss)]
        for c in self.crit_objects:
            if getattr(c, "reset_recorder", None) is not None:
                c.reset_recorder()

    def get_loss(self, results, **kwargs):
        losses = []
        for i, c in enumerate(self.crit_objects):
            if isinstance(c, CritBase):
                li, n = c(results)
            else:
                preds, gts = results[self.keys[i][0]], results[self.keys[i][1]]
                li, n = c(preds, gts), gts.size(0)
            losses.append(li * self.scales[i])
            self.loss_recorder[i].update(li.item(), n)
        return torch.stack(losses, dim=0).sum(0)

    def get_loss_info(self):
        names, info = list(self.names), [m.avg for m in self.loss_recorder]
        for c in self.crit_objects:
            if getattr(c, "get_info", None) is not None:
                n, i = c.get_info()
                names += n
                info += i
        if self.summarywriter is not None:
            self.n_current_round += 1
            for n, v in zip(names, info):
                self.summarywriter.add_scalar(n, v, global_step=self.n_current_round)
        return names, info

    def get_fieldsnames(self):
        fields, excluded = [], []
        for i, c in enumerate(self.crit_objects):
            if isinstance(c, LanguageGeneration):
                excluded.append(i)
            elif getattr(c, "get_fieldsnames", None) is not None:
                fields += c.get_fieldsnames()
        return fields + [n for i, n in enumerate(self.names) if i not in excluded]


def get_criterion(opt, summarywriter=None):
    assert isinstance(opt["crit"], list)
    objs = []
    for item in opt["crit"]:
        name = item.lower()
        if name == "lang":
            objs.append(LanguageGeneration(opt, name))
        elif name == "length":
            objs.append(nn.KLDivLoss())
        else:
            raise NotImplementedError("criterion %r (config.Constants.mapping) is not implemented" % name)
    return Criterion(objs, keys=opt["crit_key"], names=opt["crit_name"], scales=opt["crit_scale"], summarywriter=summarywriter)
